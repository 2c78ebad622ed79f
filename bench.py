#!/usr/bin/env python
"""bench.py -- perft nodes/s of the B200 path on BASELINE.json's headline workload.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU algorithm (oracle port)

Workload (N = 1 .. 8): perft(8) from the standard start position, 84 998 978 956
nodes (BASELINE.json configs[3]); one "step" is one complete perft(8).  With N > 1
(torchrun, one process per GPU) the tree is cut at ply 4 (197 281 subtrees) and
rank r walks the subtrees whose root record hashes to r (mod N); the only collective
is the u64 sum at the end of every step.

One JSON line on stdout (rank 0):
  value     nodes / device time (CUDA events on the launching stream, max over ranks)
  e2e       the same through the public API with host buffers and wall clock:
            `pabi_b200.perft(Position, depth)` at N = 1, `perft_distributed` (shard +
            ONE all-reduce per step) at N > 1
  roofline  the dominant kernel (K2, the fused last two plies) against the roofline that
            binds it, integer instruction issue; SURVEY.md 8(d)'s byte model beside it
  other_configs   BASELINE.json configs 1, 2, 3, 5, measured outside the headline's timed
            region, every result checked (known answers / tests/golden fixtures)
  cpu_baseline    the oracle port of pabi's perft on the host cores: all threads and one
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

DEPTH = 8
NODES = {5: 4_865_609, 6: 119_060_324, 7: 3_195_901_860, 8: 84_998_978_956}
# sum_{i=1}^{d-1} perft(i): every node of plies 1..d-1 is written once and read once at 64 B
INTERIOR = {6: 5_072_212, 7: 124_132_536, 8: 3_320_034_396}
SPLIT_PLY = 4
METRIC = "perft nodes/sec"
UNIT = "nodes/s"
SM_COUNT, SCHEDULERS_PER_SM = 148, 4


def measured_peaks() -> dict:
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        m = json.loads(f.read_text())
        return {"hbm_gbs": m["hbm_gbs"], "sm_max_mhz": m.get("sm_max_mhz", 1965.0), "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for r in self.rows if len(r) >= 7 and r[0].isdigit()]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in rows)]
        return {"sm_mhz": statistics.median(int(r[0]) for r in rows), "sm_max_mhz": int(rows[0][1]),
                "power_w_max": max(float(r[2]) for r in rows), "samples": len(rows), "reasons": reasons}


# ---- the CPU legs: the only places where bench.py executes oracle/ ------------------------------------------

def cpu_perft(depth: int, threads: int):
    """The CPU checker (oracle/, a port of the reference's algorithm) timed on the host cores."""
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle
    p = oracle.starting()
    t0 = time.perf_counter()
    n = oracle.perft(p, depth, threads)
    dt = time.perf_counter() - t0
    assert n == NODES[depth], (n, NODES[depth])
    return n, dt


PORT_NOTE = ("oracle/ C port of pabi's perft (copy-make with incremental Zobrist updates, depth-1 bulk count, PEXT tables); "
             "the Rust reference cannot be built here (no cargo/rustc)")


def single_thread_sample() -> dict:
    """What pabi's perft is (src/chess/position.rs:909-924, single-threaded): perft(6), and perft(7) when it fits ~15 s."""
    n, dt = cpu_perft(6, 1)
    depth = 6
    if dt * 27 < 16:
        n, dt = cpu_perft(7, 1)
        depth = 7
    return {"value": n / dt, "unit": UNIT, "cores": 1, "sample": f"startpos perft({depth}) ({n} nodes, {dt:.2f} s) on one thread"}


def cpu_baseline_sample(threads: int) -> dict:
    """~10-30 s of CPU work on the same tree: all host threads (perft(7) if it fits), then one thread."""
    n, dt = cpu_perft(6, threads)
    depth = 6
    if dt * 27 < 40:
        n, dt = cpu_perft(7, threads)
        depth = 7
    return {"value": n / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"startpos perft({depth}) ({n} nodes, {dt:.2f} s), root split at ply 2 over {threads} threads; " + PORT_NOTE,
            "single_thread": single_thread_sample()}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    times, nodes = [], 0
    sample_depth = 7
    _, gauge = cpu_perft(6, threads)  # sizes the sample: perft(7) is ~27x perft(6)
    if gauge * 27 * (args.warmup + args.steps) > 240:
        sample_depth = 6
    for i in range(args.warmup + args.steps):
        n, dt = cpu_perft(sample_depth, threads)
        if i >= args.warmup:
            times.append(dt)
            nodes += n
    total = sum(times)
    value = nodes / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"startpos perft({DEPTH}), {NODES[DEPTH]} nodes",
                   "sample": f"each step = startpos perft({sample_depth}) ({NODES[sample_depth]} nodes) on all {threads} host threads; "
                             "a rate, not the same tree size: the denominator of any GPU/CPU ratio grows with the host's core count"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"startpos perft({sample_depth}) x {len(times)} steps, {threads} threads; " + PORT_NOTE,
                         "single_thread": single_thread_sample()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- the roofline of the dominant kernel ---------------------------------------------------------------------

KERNEL_SOURCES = ("chess.cuh", "incremental.cuh", "kernels.cuh")


def kernel_source_digest() -> str:
    h = hashlib.sha256()
    for name in KERNEL_SOURCES:
        h.update((ROOT / "pabi_b200" / "csrc" / name).read_bytes())
    return h.hexdigest()[:16]


def leaf_roofline(depth, world, steps, leaf_ms, leaf_launches, leaf_parents, leaf_children, device_ms, device_ms_max, clocks, peaks):
    """K2 (`k_tile<TILE_LEAF>`, the last two plies fused) against integer instruction issue: the kernel keeps its state
    in registers and shared memory, so what bounds it is the rate at which the four schedulers of an SM issue warp
    instructions (regime ii), not HBM.  The instruction counts come from `ncu` on the same launches
    (profiles/leaf_kernel_counts.json, written by tools/save_round_r02.sh from an ncu pass over one perft(8)); the
    duration and the clock are measured live."""
    sm_mhz = (clocks or {}).get("sm_mhz") or peaks["sm_max_mhz"]
    peak_issue = SM_COUNT * SCHEDULERS_PER_SM * sm_mhz * 1e6 / 1e9         # G warp-instructions / s
    out = {"bound": "int-issue", "kernel": "k_tile<TILE_LEAF> (K2: the last two plies fused, children never written)",
           "unit": "Gwarp-inst/s", "peak": peak_issue,
           "peak_source": f"{SM_COUNT} SMs x {SCHEDULERS_PER_SM} schedulers x {sm_mhz:.0f} MHz (median SM clock sampled during the timed region)",
           "launches": leaf_launches, "avg_launch_ms": leaf_ms / max(1, leaf_launches),
           "share_of_step": leaf_ms / device_ms if device_ms else None}
    counts_file = ROOT / "profiles" / "leaf_kernel_counts.json"
    if counts_file.exists() and leaf_children:
        c = json.loads(counts_file.read_text())
        per_step_children = leaf_children / steps
        exact = world == 1 and depth == 8 and per_step_children == c["leaf_children"]
        scale = per_step_children / c["leaf_children"]
        warp_inst = c["warp_instructions"] * scale * steps
        achieved = warp_inst / (leaf_ms * 1e-3) / 1e9
        alu_peak = peak_issue * 0.5  # the ALU pipe (LOP3 / SHF / IADD3 / SEL / ISETP on 32-bit halves) issues every other cycle
        out.update({
            "achieved": achieved, "frac": achieved / peak_issue,
            "warp_instructions_per_launch": warp_inst / max(1, leaf_launches),
            "instructions": {"source": c["source"], "exact_for_this_workload": exact,
                             "scaled_by_children": None if exact else scale,
                             "kernel_source_digest": c.get("kernel_source_digest"),
                             "matches_current_source": c.get("kernel_source_digest") == kernel_source_digest(),
                             "thread_inst_per_warp_inst": c["thread_instructions"] / c["warp_instructions"],
                             "warp_inst_per_child": c["warp_instructions"] / c["leaf_children"]},
            "alu_pipe": {"achieved": c["alu_instructions"] * scale * steps / (leaf_ms * 1e-3) / 1e9, "peak": alu_peak, "unit": "Gwarp-inst/s",
                         "frac": c["alu_instructions"] * scale * steps / (leaf_ms * 1e-3) / 1e9 / alu_peak,
                         "note": "the binding pipe: 64-bit bitboard logic on a 32-bit datapath"},
            "traffic": c["dram_bytes_per_parent"] * leaf_parents / max(1, leaf_launches) if c.get("dram_bytes_per_parent") else None,
            "traffic_note": (f"{c['dram_bytes_per_parent']:.1f} B/parent (dram__bytes_read+write of the ncu --set full capture of the "
                             "depth-7 launch) x parents per launch: the kernel reads its parents and writes nothing")
                            if c.get("dram_bytes_per_parent") else None,
            "from_ncu_capture": c.get("capture"),
        })
    else:
        out.update({"achieved": None, "frac": None, "traffic": None,
                    "note": "profiles/leaf_kernel_counts.json is missing: no instruction counts for the issue roofline"})
    # SURVEY.md 8(d)'s shared byte figure, kept beside it: 64 B read per parent, 64 B written + 64 B read per child
    leaf_bytes = 64 * leaf_parents + 128 * leaf_children
    hbm = leaf_bytes / (leaf_ms * 1e-3) / 1e9 if leaf_ms else 0.0
    out["hbm_model"] = {
        "algorithmic_bytes_per_launch": leaf_bytes / max(1, leaf_launches), "achieved": hbm, "peak": peaks["hbm_gbs"], "unit": "GB/s",
        "frac": hbm / peaks["hbm_gbs"], "peak_source": peaks["source"],
        "whole_step": {"algorithmic_bytes": 128 * INTERIOR.get(depth, 0),
                       "achieved": 128 * INTERIOR.get(depth, 0) * steps / (device_ms_max * 1e-3) / 1e9},
        "note": ("not a physical fraction: the byte model has every node of plies 1..d-1 written to and read from HBM once, but the "
                 "fused kernel never materialises the last ply but one and settles most of its moves per parent; the kernel's real "
                 "DRAM bytes are `traffic`")}
    return out


# ---- BASELINE.json's other configs (outside the headline's timed region) --------------------------------------

def sha(arr) -> str:
    return hashlib.sha256(arr.tobytes()).hexdigest()


def best_of(engine, fn, reps):
    out = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        out.append((engine.last_timing.device_ms, (time.perf_counter() - t0) * 1e3, r))
    return min(out, key=lambda x: x[0])


def other_configs_single(pb, engine, peaks) -> dict:
    """configs[0], [1], [4] on one GPU (rank 0).  Fractions are of the HBM roofline with SURVEY.md 8(d)'s algorithmic bytes."""
    import numpy as np
    from pabi_b200 import workloads as W
    hbm = peaks["hbm_gbs"] * 1e9
    fixtures = json.loads((ROOT / "tests" / "golden" / "workload_fixtures.json").read_text())
    out = {}
    # configs[0]: startpos perft(6), the reference's own bench case (benches/chess.rs:54)
    dev, wall, n = best_of(engine, lambda: engine.perft(pb.Position.starting(), 6), 7)
    assert n == NODES[6]
    out["perft6_startpos"] = {"nodes": n, "device_ms": dev, "e2e_ms": wall, "value": n / dev * 1e3, "e2e_value": n / wall * 1e3, "unit": UNIT,
                              "launches": engine.last_timing.kernel_launches,
                              "roofline": {"bound": "hbm", "algorithmic_bytes": 128 * INTERIOR[6], "frac": 128 * INTERIOR[6] / (dev * 1e-3) / hbm,
                                           "note": "launch-latency bound: six plies in a fraction of a millisecond"}}
    # configs[1]: the reference's perft suite (tests/chess.rs:622-818, benches/chess.rs:52-86), known answers
    vec = json.loads((ROOT / "tests" / "golden" / "reference_vectors.json").read_text())
    cases = vec["perft"] + vec["bench_perft"]
    positions = [pb.Position.starting() if c["fen"] == "startpos" else pb.Position.try_from(c["fen"]) for c in cases]
    tot_dev = tot_wall = 0.0
    for c, p in zip(cases, positions):
        dev, wall, n = best_of(engine, lambda: engine.perft(p, c["depth"]), 3)
        assert n == c["nodes"], c
        tot_dev += dev
        tot_wall += wall
    nodes = sum(c["nodes"] for c in cases)
    dev_b, wall_b, counts = best_of(engine, lambda: engine.perft_batch_depths(positions, [c["depth"] for c in cases]), 3)
    assert counts.tolist() == [c["nodes"] for c in cases]
    out["perft_suite"] = {"vectors": len(cases), "nodes": nodes, "one_root_at_a_time": {"device_ms": tot_dev, "e2e_ms": tot_wall},
                          "one_batch_per_root_depths": {"device_ms": dev_b, "e2e_ms": wall_b, "api": "pabi_cuda_perft_batch_depths"},
                          "value": nodes / min(tot_dev, dev_b) * 1e3, "unit": UNIT}
    # configs[4]: ordered move lists of 1 M playout positions, produced on the device, lists checked by digest
    positions = W.playout_positions(engine, 1_000_000)
    fx = fixtures["playouts"]
    assert len(positions) == fx["positions"] and sha(positions) == fx["sha256"], "playout positions differ from the oracle's"
    _, wall, (offsets, moves) = best_of(engine, lambda: engine.generate_moves_batch(positions), 3)
    assert len(moves) == fx["moves"] and sha(offsets) == fx["offsets_sha256"] and sha(moves) == fx["moves_sha256"], "move lists differ"
    pinned_in = pb.pinned_empty(len(positions), positions.dtype)
    pinned_in[:] = positions
    pinned_offsets = pb.pinned_empty(len(positions) + 1, np.uint32)
    pinned_moves = pb.pinned_empty(40 * len(positions), np.uint16)
    _, wall_p, (offsets, moves) = best_of(engine, lambda: engine.generate_moves_batch(pinned_in, pinned_offsets, pinned_moves), 5)
    assert sha(offsets) == fx["offsets_sha256"] and sha(moves) == fx["moves_sha256"]
    chunks = engine.last_timing.kernel_launches
    # device time with the records resident in HBM (64-byte SoA records, outputs on the device): one launch of the fused kernel
    import torch
    n = len(positions)
    rec = torch.empty(8 * n, dtype=torch.int64, device="cuda")
    d_off = torch.empty(n + 1, dtype=torch.int32, device="cuda")
    d_mv = torch.empty(int(len(moves)) + 1024, dtype=torch.int16, device="cuda")
    engine.pack_records(positions, rec.data_ptr(), n)
    dev = min(engine.generate_moves_resident(rec.data_ptr(), n, n, d_off.data_ptr(), d_mv.data_ptr(), d_mv.numel())[1].device_ms for _ in range(7))
    t = engine.last_timing
    assert sha(d_off.cpu().numpy().view(np.uint32)) == fx["offsets_sha256"] and sha(d_mv[:len(moves)].cpu().numpy().view(np.uint16)) == fx["moves_sha256"]
    out["move_lists_1m"] = {"positions": n, "moves": int(len(moves)), "device_ms": dev, "e2e_ms_pageable": wall,
                            "e2e_ms_pinned": wall_p, "h2d_bytes": int(positions.nbytes), "d2h_bytes": int(offsets.nbytes + moves.nbytes),
                            "value": n / dev * 1e3, "e2e_value": n / wall_p * 1e3, "unit": "positions/s",
                            "launches": t.kernel_launches, "e2e_chunks": chunks,
                            "api": "device: pabi_cuda_generate_moves_device on resident records; e2e: pabi_cuda_generate_moves_batch, host arrays, "
                                   "copies of a chunk overlapping the kernel of the next",
                            "checked": "sha256 of offsets and moves == the CPU oracle's (tests/golden/workload_fixtures.json), both paths",
                            "roofline": {"bound": "hbm", "algorithmic_bytes": t.algorithmic_bytes,
                                         "frac": t.algorithmic_bytes / (dev * 1e-3) / hbm,
                                         "note": "issue-bound in generate_moves (unrelated positions diverge), far from the byte roofline"}}
    return out


def book_config(pb, engine, world, device, peaks) -> dict:
    """configs[2]: perft(4) over every position of the book, position i on rank i mod N, counts gathered over NCCL and checked
    position by position.  The real book is absent from the mount (.MISSING_LARGE_BLOBS), so this is the documented stand-in."""
    import torch.distributed as dist
    from pabi_b200 import distributed as D
    from pabi_b200 import workloads as W
    fx = json.loads((ROOT / "tests" / "golden" / "workload_fixtures.json").read_text())["book"]
    book = W.chess324_synth_book(engine)
    assert len(book) == fx["positions"] and sha(book) == fx["sha256"], "book positions differ from the oracle's"
    best_dev = best_wall = None
    for _ in range(4):
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        counts = D.perft_batch_distributed(engine, book, fx["depth"], device)
        wall = (time.perf_counter() - t0) * 1e3
        dev = engine.last_timing.device_ms
        dev, wall = D.all_reduce_max([dev, wall], device)
        assert counts.tolist() == fx["perft"], "per-position counts differ"
        if best_dev is None or dev < best_dev:
            best_dev, best_wall = dev, wall
    nodes = fx["nodes"]
    return {"book": "chess324-synth (stand-in: the reference's Chess324 book is absent from the mount)", "positions": len(book), "depth": fx["depth"],
            "nodes": nodes, "n_gpus": world, "device_ms": best_dev, "e2e_ms": best_wall, "value": nodes / best_dev * 1e3, "unit": UNIT,
            "checked": "every per-position count == the CPU oracle's (tests/golden/workload_fixtures.json)",
            "sharding": "position i on rank i mod N, u64 counts gathered by one all-reduce over disjoint supports"}


# ---- the headline ---------------------------------------------------------------------------------------------

def run_gpu(args) -> None:
    import torch
    import torch.distributed as dist

    import pabi_b200 as pb
    from pabi_b200 import distributed as D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line: anything a library prints there (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
        cpu_group = dist.new_group(backend="gloo")  # host-side waits that must not put a spinning kernel on the GPUs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    engine = pb.Engine(local, args.capacity)
    root = pb.Position.starting()
    depth = args.depth
    expected = NODES.get(depth)

    # single-process multi-GPU (what a Rust host would call): checked once over distinct devices before anything is timed
    multi = None
    if world > 1 and not args.no_extra:
        if rank == 0 and torch.cuda.device_count() >= world:
            engines = [engine] + [pb.Engine(d, args.capacity) for d in range(world) if d != local]
            nodes, _ = pb.perft_multi(engines, root, 7, SPLIT_PLY)
            assert nodes == NODES[7], f"perft_multi over {world} devices: {nodes}"
            times = []
            for _ in range(4):
                t0 = time.perf_counter()
                nodes, per_ctx = pb.perft_multi(engines, root, depth, SPLIT_PLY)
                times.append(((time.perf_counter() - t0) * 1e3, max(per_ctx)))
                assert expected is None or nodes == expected
            wall, dev = min(times)
            multi = {"api": "pabi_cuda_perft_multi (one process, one host thread and context per device, host-side sum)", "devices": world,
                     "e2e_ms": wall, "device_ms_max": dev, "value": (expected or nodes) / dev * 1e3, "e2e_value": (expected or nodes) / wall * 1e3,
                     "unit": UNIT, "checked": "perft(7) and every timed perft exact over distinct devices"}
            for e in engines[1:]:
                e.close()
        dist.barrier(group=cpu_group)  # the other ranks wait on the host: rank 0 is using their devices
        barrier()

    def step():
        """One perft through the public API; at N > 1 the shard walk AND the one collective of the path."""
        if world == 1:
            n = pb.perft(root, depth, engine)
        else:
            n = D.perft_distributed(engine, root, depth, SPLIT_PLY, device)
        return n, engine.last_timing

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
    wall0 = time.perf_counter()
    device_ms = leaf_ms = 0.0
    launches = leaf_launches = leaf_parents = leaf_children = 0
    for _ in range(args.steps):
        n, t = step()
        if expected is not None and n != expected:
            raise SystemExit(f"wrong node count: {n} != {expected}")
        device_ms += t.device_ms
        leaf_ms += t.leaf_kernel_ms
        launches += t.kernel_launches
        leaf_launches += t.leaf_kernel_launches
        leaf_parents += t.leaf_parents
        leaf_children += t.leaf_children
    barrier()
    wall_s = time.perf_counter() - wall0
    clocks = sampler.stop() if sampler else None
    device_ms_max, wall_ms_max, leaf_ms_max = D.all_reduce_max([device_ms, wall_s * 1e3, leaf_ms], device)
    launches_total = D.all_reduce_sum_u64(launches, device)
    nodes_total = (expected if expected is not None else n) * args.steps

    extra = {}
    if not args.no_extra and depth == DEPTH:
        peaks = measured_peaks()
        book = book_config(pb, engine, world, device, peaks)  # every rank takes part
        if rank == 0:
            extra["perft4_book"] = book
            if world == 1:
                extra.update(other_configs_single(pb, engine, peaks))
            if multi:
                extra["perft8_single_process_multi_gpu"] = multi

    if rank == 0:
        peaks = measured_peaks()
        line = {
            "metric": METRIC, "value": nodes_total / (device_ms_max * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": device_ms_max / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"startpos perft({depth}), {NODES.get(depth, '?')} nodes, bit-exact count checked every step",
                       "sharding": "single GPU" if world == 1 else f"root subtrees at ply {SPLIT_PLY}, cut by a hash of the subtree's root record (rank = hash mod N); one u64 all-reduce (NCCL) per step",
                       "l2": "frontier buffers (>= 7.6 GB at ply 6) are far larger than the 126 MB L2; no flush needed",
                       "frontier_capacity_records": args.capacity or 32 << 20},
            "e2e": {"value": nodes_total / (wall_ms_max * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 112,
                    "d2h_bytes_per_step": 16, "ms_per_step": wall_ms_max / args.steps,
                    "api": "pabi_b200.perft(Position, depth) -> pabi_cuda_perft" if world == 1 else
                           "pabi_b200.distributed.perft_distributed -> pabi_cuda_perft_shard + one all_reduce per step"},
            "gpu_launches": launches_total,
            "clocks": clocks,
            "roofline": leaf_roofline(depth, world, args.steps, leaf_ms, leaf_launches, leaf_parents, leaf_children, device_ms,
                                      device_ms_max, clocks, peaks),
        }
        if extra:
            line["other_configs"] = extra
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline_sample(os.cpu_count() or 1)
        print(json.dumps(line), file=json_out, flush=True)
    engine.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="pabi_b200", choices=["pabi_b200", "reference"])
    ap.add_argument("--depth", type=int, default=DEPTH)
    ap.add_argument("--capacity", type=int, default=0, help="frontier records per level buffer (0 = library default)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the other_configs legs (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
