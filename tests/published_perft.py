"""Published perft values from outside the reference (the well-known edge-case suite circulated on
the Chess Programming Wiki / TalkChess: en passant legality, castling through and out of attack,
promotions, under-promotion checks, stalemate/checkmate corners, and the 218-move position).
They pin the oracle against independent implementations, and the GPU path against both."""

PUBLISHED = [
    # (fen, depth, nodes)
    ("3k4/3p4/8/K1P4r/8/8/8/8 b - - 0 1", 6, 1134888),        # illegal ep move #1
    ("8/8/4k3/8/2p5/8/B2P2K1/8 w - - 0 1", 6, 1015133),       # illegal ep move #2
    ("8/8/1k6/2b5/2pP4/8/5K2/8 b - d3 0 1", 6, 1440467),      # ep capture checks opponent
    ("5k2/8/8/8/8/8/8/4K2R w K - 0 1", 6, 661072),            # short castling gives check
    ("3k4/8/8/8/8/8/8/R3K3 w Q - 0 1", 6, 803711),            # long castling gives check
    ("r3k2r/1b4bq/8/8/8/8/7B/R3K2R w KQkq - 0 1", 4, 1274206),  # castle rights
    ("r3k2r/8/3Q4/8/8/5q2/8/R3K2R b KQkq - 0 1", 4, 1720476),   # castling prevented
    ("2K2r2/4P3/8/8/8/8/8/3k4 w - - 0 1", 6, 3821001),        # promote out of check
    ("8/8/1P2K3/8/2n5/1q6/8/5k2 b - - 0 1", 5, 1004658),      # discovered check
    ("4k3/1P6/8/8/8/8/K7/8 w - - 0 1", 6, 217342),            # promote to give check
    ("8/P1k5/K7/8/8/8/8/8 w - - 0 1", 6, 92683),              # under promote to give check
    ("K1k5/8/P7/8/8/8/8/8 w - - 0 1", 6, 2217),               # self stalemate
    ("8/k1P5/8/1K6/8/8/8/8 w - - 0 1", 7, 567584),            # stalemate & checkmate
    ("8/8/2k5/5q2/5n2/8/5K2/8 b - - 0 1", 4, 23527),          # stalemate & checkmate
    # CPW "position 5" and "position 6" (also in the reference), one ply deeper than the reference pins
    ("rnbq1k1r/pp1Pbppp/2p5/8/2B5/8/PPP1NnPP/RNBQK2R w KQ - 1 8", 4, 2103487),
    # startpos beyond the reference's own depth
    ("rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1", 7, 3195901860),
]

# the position with the most legal moves known (218)
MAX_MOVES_FEN = "R6R/3Q4/1Q4Q1/4Q3/2Q4Q/Q4Q2/pp1Q4/kBNN1KB1 w - - 0 1"
