#!/usr/bin/env python3
"""Benchmark of the hot path: one step = one pass of a query operator over a 10k-query batch on the
synthetic 10M-doc / 1M-term Zipfian index (BASELINE.json configs[3]/[4]), index resident in HBM.

  python bench.py --gpus N --steps K --warmup W [--op ranked_and|wand|maxscore|and] [--scaling weak|strong]
  python bench.py --impl reference ...     times the compiled reference (oracle/_ref) on the host cores

value   queries/s with the batch already resident in HBM (CUDA events around the K steps, max over ranks)
e2e     queries/s through ds2i_gpu_query_batch with HOST buffers (H2D of the queries and D2H of
        counts + top-k inside the timed region, host-side query preparation included)
roofline  algorithmic bytes (device counters = blocks/bytes the reference algorithm decodes)
          / average kernel time, against MEASURED_PEAKS.json hbm_gbs
cpu_baseline  the reference's own operators (oracle/_ref/ref_tool, compiled from /root/reference) on all
          host cores over the same 10k queries, same index file; plus the single-thread op_perftest protocol
          (queries.cpp:13-62: 3 passes, first discarded) on a 1000-query prefix.
parity  every query of the step against the reference's own results (-ffp-contract=off build), per operator.
also    wand, maxscore, or, ranked_or (same protocol); decode = BASELINE config 2 (batched decode of all 1M lists, block_optpfor and
        block_interpolative, checksum of all postings + bit-exact sample vs the reference, reference scan on 1 thread and all
        cores beside it); pef = config 3 (`opt` index of the same collection: full scans and next_geq sweeps); opt = the
        query operators over the `opt` index.  N = 1 only.
Multi-GPU: index replicated per GPU, queries dealt to the ranks cost-balanced; every rank evaluates its shard and the fused
per-shard results (counts + top-k scores + docids) are gathered with ONE NCCL all_gather enqueued behind the kernels (no host
synchronisation inside a step).  scaling "weak": 10k queries per GPU per step; also.strong: 10k queries in total.
"""
import argparse
import json
import os
import statistics
import struct
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "queries/sec (%s, %s, synthetic 10M-doc/1M-term Zipfian index, 10k-query batch, top-10)"
WORKLOAD = "ranked top-10 over synthetic Zipfian index (configs[3]/[4])"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def data_dir(args):
    base = os.environ.get("DS2I_BENCH_DATA", "/tmp/ds2i_b200_data")
    return os.path.join(base, "S_%d_%d_%d_q%d" % (args.docs, args.terms, args.seed, args.queries_total))


def ensure_data(args, rank=0, types=("block_optpfor",)):
    """Synthetic collection -> ds2i-format indexes + wand data + queries, built by OUR builder
    (ds2i_b200/csrc/builder.cpp; byte-identical to the reference's create_freq_index output)."""
    d = data_dir(args)
    paths = {"wand": os.path.join(d, "S.wand"), "queries": os.path.join(d, "S.queries")}
    for t in types:
        paths[t] = os.path.join(d, "S.%s.idx" % t)
    paths["index"] = paths.get("block_optpfor")
    missing = [t for t in types if not os.path.exists(os.path.join(d, "DONE." + t))]
    if missing:
        if rank == 0:
            from ds2i_b200 import build
            build.build()
            os.makedirs(d, exist_ok=True)
            t0 = time.time()
            subprocess.run([build.BUILDER, "synth", os.path.join(d, "S"), str(args.docs), str(args.terms), str(args.seed), "0",
                            str(args.queries_total), ":".join(missing)], check=True)
            log("[bench] built synthetic %s in %.1f s -> %s" % (",".join(missing), time.time() - t0, d))
            for t in missing:
                open(os.path.join(d, "DONE." + t), "w").write("ok\n")
        else:
            while any(not os.path.exists(os.path.join(d, "DONE." + t)) for t in types):
                time.sleep(0.5)
    return paths


class ClockSampler:
    """SM clock and throttle reasons sampled every few ms through NVML while the timed region runs."""

    def __init__(self, gpu_index):
        self.samples = []
        self.stop = False
        self.gpu = gpu_index
        self.t = threading.Thread(target=self.run, daemon=True)
        self.max_mhz = None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.gpu)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            while not self.stop:
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    reasons = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    reasons = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self.samples.append((sm, reasons))
                time.sleep(0.004)
        except Exception:          # NVML missing: fall back to one nvidia-smi query
            try:
                r = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"],
                                   stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5)
                a, b = [float(x) for x in r.stdout.strip().split(",")]
                self.samples.append((a, 0)); self.max_mhz = b
            except Exception:
                pass

    def __enter__(self):
        if os.environ.get("DS2I_BENCH_NO_CLOCKS"):        # diagnostic: is the NVML polling itself in the way?
            return self
        self.t.start()
        time.sleep(0.05)
        return self

    def __exit__(self, *a):
        self.stop = True
        if self.t.is_alive():
            self.t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        reasons = [n for n, b in bits.items() if any(s[1] & b for s in self.samples)]
        return {"sm_mhz": statistics.median(s[0] for s in self.samples), "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(self.samples)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def traffic_per_launch(op):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/traffic.json): a
    STATIC figure of that capture, not measured in this run (a run under ncu is never a bench value)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        return json.load(open(p)).get(op)
    return None


def ref_tool(*args, strict=False, timeout=900):
    tool = os.path.join(ROOT, "oracle", "_ref", "ref_tool_strict" if strict else "ref_tool")
    if not os.access(tool, os.X_OK):
        raise RuntimeError("oracle/_ref/ref_tool missing (built in the build container by oracle/Makefile)")
    r = subprocess.run([tool] + [str(a) for a in args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, check=True, timeout=timeout)
    return r.stdout


def run_reference_tool(paths, op, threads, sample, passes, itype="block_optpfor"):
    return json.loads(ref_tool("bench", itype, paths[itype], paths["wand"], paths["queries"], op, threads, sample, passes).strip().splitlines()[-1])


def reference_block_profile(paths, op, sample):
    """Bytes the REFERENCE algorithm decodes per query: ds2i's own block_profiler (block_profiler.hpp:40-54,
    what profile_queries.cpp uses) run over the first `sample` queries by oracle/_ref/ref_tool.  B_q = docs payload of the
    blocks decoded at least once + freqs payload likewise + 8 B (block_max + endpoint) per decoded docs block."""
    import numpy as np
    out = os.path.join(os.path.dirname(paths["index"]), "profile.%s.bin" % op)
    ref_tool("profile", "block_optpfor", paths["index"], paths["wand"], paths["queries"], op, out, sample, timeout=300)
    raw = np.fromfile(out, dtype="<u8")
    nq = int(raw[0])
    a = raw[1:1 + 6 * nq].reshape(nq, 6).astype(np.float64)
    return {"queries": nq, "docs_blocks_per_query": float(a[:, 0].mean()), "freqs_blocks_per_query": float(a[:, 1].mean()),
            "bytes_per_query": float((a[:, 2] + a[:, 3] + 8 * a[:, 0]).mean()), "list_bytes_per_query": float(a[:, 4].mean())}


def cpu_baseline_for(paths, op, nq, itype="block_optpfor", single_prefix=1000):
    """The reference's own operator on the host: all cores over the whole batch (thread t takes queries t, t+n, ...:
    profile_queries.cpp:21-39), and the single-thread op_perftest protocol (queries.cpp:13-62: 3 passes, first discarded)
    on a bounded prefix."""
    cores = os.cpu_count() or 1
    out = run_reference_tool(paths, op, cores, nq, 2, itype)
    base = {"value": nq / out["pass_seconds"][-1], "unit": "queries/s", "cores": cores, "kind": "reference",
            "sample": "all %d queries of the step, second of 2 passes, ds2i %s_query compiled from the reference sources (-O3 -march=x86-64-v3), "
                      "thread t takes queries t,t+n,.. (profile_queries.cpp:21-39)" % (nq, op)}
    n1 = min(single_prefix, nq)
    o1 = run_reference_tool(paths, op, 1, n1, 3, itype)
    secs = o1["pass_seconds"][1:]
    base["single_thread"] = {"value": n1 / (sum(secs) / len(secs)), "unit": "queries/s", "cores": 1, "avg_us_per_query": 1e6 * (sum(secs) / len(secs)) / n1,
                             "sample": "first %d queries, op_perftest protocol (queries.cpp:13-62): 3 passes, first discarded" % n1}
    return base


def parity_block(d, paths, op, k, qidx, counts, scores, itype="block_optpfor"):
    """Every query of the step against the reference's own results for the same index file.  qidx[j] = position in the
    query file of the j-th query of the step (the step runs the queries in the scheduler's cost-balanced order)."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import load_dump
    tmp = os.path.join(os.path.dirname(paths["index"]), "check.%s.%s.bin" % (itype, op))
    nq = len(qidx)
    ref_tool("dump", itype, paths[itype], paths["wand"], paths["queries"], tmp, op, k, int(max(qidx)) + 1 if nq else 0, strict=True)
    ec, es = load_dump(tmp, (op,))[op]
    ec, es = ec[qidx], es[qidx]
    out = {"queries_checked": int(nq), "counts_bit_exact": bool(np.array_equal(ec, counts[:nq])), "against": "ds2i reference, -ffp-contract=off build, same index file"}
    if op in d.RANKED:
        out["scores_bit_exact"] = bool(np.array_equal(es.view(np.uint32), scores[:nq].view(np.uint32)))
        out["scores_max_rel_err"] = float(np.max(np.abs(es.astype(np.float64) - scores[:nq]) / np.maximum(np.abs(es), 1e-30))) if nq else 0.0
        out["tolerance"] = 1e-5
        out["ok"] = out["counts_bit_exact"] and out["scores_max_rel_err"] <= 1e-5
    else:
        out["ok"] = out["counts_bit_exact"]
    return out


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    paths = ensure_data(args)
    cores = os.cpu_count() or 1
    sample = args.queries if args.ref_sample <= 0 else min(args.ref_sample, args.queries)
    out = run_reference_tool(paths, args.op, cores, sample, args.warmup + args.steps)
    secs = out["pass_seconds"][args.warmup:]
    per_step = sum(secs) / len(secs)
    qps = sample / per_step
    line = {
        "impl": "reference", "metric": METRIC % (args.op, "block_optpfor"), "value": qps, "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32+f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "op": args.op, "index_type": "block_optpfor",
                   "num_docs": args.docs, "num_terms": args.terms, "queries_per_step": sample, "k": args.k, "seed": args.seed},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "reference",
                         "sample": "%d of the %d queries per step, ds2i %s_query compiled from /root/reference (-O3 -march=x86-64-v3; the GPU box has no "
                                   "/root/reference to rebuild with -march=native), thread t takes queries t,t+n,.. (profile_queries.cpp:21-39)" % (sample, args.queries, args.op)},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def decode_leg(d, args, paths, itype, peak, steps, warmup):
    """BASELINE config 2: batched decode of every list of the index, outputs materialised in HBM."""
    import numpy as np
    idx = d.Index(paths[itype], itype, 0)
    terms = np.arange(idx.size(), dtype=np.uint32)
    for _ in range(warmup):
        idx.decode_lists_checksum(terms)
    ms = []
    for _ in range(steps):
        postings, sd, sf, m = idx.decode_lists_checksum(terms)
        ms.append(m)
    m = sum(ms) / len(ms)
    in_bytes = os.path.getsize(paths[itype])
    out_bytes = 8 * postings
    cores = os.cpu_count() or 1
    leg = {"metric": "decoded ints/sec (batched %s block decode, all %d lists of the synthetic index, docids + freqs)" % (itype, idx.size()),
           "value": 2 * postings / (m * 1e-3), "unit": "ints/s", "ms_per_step": m, "postings": postings, "gpu_launches": 2 * steps,
           "roofline": {"bound": "hbm", "achieved": (in_bytes + out_bytes) / (m * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": (in_bytes + out_bytes) / (m * 1e-3) / 1e9 / peak, "traffic": None,
                        "algorithmic_bytes_per_launch": in_bytes + out_bytes, "bytes_in": in_bytes, "bytes_out": out_bytes,
                        "note": "in = the index file (compressed blocks + block_max / endpoint arrays), out = 4 B docid + 4 B freq per posting"}}
    # parity 1: checksum of ALL postings against the reference's own scan (which is also the all-cores CPU baseline)
    ref_all = json.loads(ref_tool("scan", itype, paths[itype], "all", cores, 1).strip().splitlines()[-1])
    ref_one = json.loads(ref_tool("scan", itype, paths[itype], "first:3000", 1, 1).strip().splitlines()[-1])
    leg["cpu_baseline"] = {"value": 2 * ref_all["postings_per_s"], "unit": "ints/s", "cores": cores, "kind": "reference",
                           "sample": "next()/docid()/freq() scan of all %d lists (%d postings), list i -> thread i %% n" % (ref_all["lists"], ref_all["postings"]),
                           "single_thread": {"value": 2 * ref_one["postings_per_s"], "unit": "ints/s", "cores": 1,
                                             "sample": "the 3000 longest lists (%d postings)" % ref_one["postings"]}}
    par = {"all_postings": postings, "sum_docids_equal": sd == ref_all["sum_docids"], "sum_freqs_equal": sf == ref_all["sum_freqs"],
           "postings_equal": postings == ref_all["postings"]}
    # parity 2: >= 10k sampled lists bit for bit
    rng = np.random.default_rng(20261017)
    sample = np.unique(np.concatenate([rng.integers(0, idx.size(), size=10000), np.arange(64)])).astype(np.uint32)
    tfile = os.path.join(os.path.dirname(paths[itype]), "sample_terms.txt")
    open(tfile, "w").write("\n".join(str(int(t)) for t in sample) + "\n")
    lfile = os.path.join(os.path.dirname(paths[itype]), "sample_lists.%s.bin" % itype)
    ref_tool("lists", itype, paths[itype], tfile, lfile)
    offs, docs, freqs, _ = idx.decode_lists(sample)
    raw = np.fromfile(lfile, dtype=np.uint32)
    pos, ok = 0, True
    for i in range(len(sample)):
        n = int(raw[pos]) | (int(raw[pos + 1]) << 32)
        pos += 2
        a, b = int(offs[i]), int(offs[i + 1])
        ok = ok and n == b - a and np.array_equal(raw[pos:pos + n], docs[a:b]) and np.array_equal(raw[pos + n:pos + 2 * n], freqs[a:b])
        pos += 2 * n
    par["sampled_lists"] = int(len(sample)); par["sampled_postings"] = int(offs[-1]); par["sampled_lists_bit_exact"] = bool(ok)
    par["ok"] = bool(ok and par["sum_docids_equal"] and par["sum_freqs_equal"] and par["postings_equal"])
    leg["parity"] = par
    idx.close()
    return leg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--op", default="ranked_and")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--docs", type=int, default=10_000_000)
    ap.add_argument("--terms", type=int, default=1_000_000)
    ap.add_argument("--queries", type=int, default=10_000, help="queries per step (per GPU under weak scaling, in total under strong scaling)")
    ap.add_argument("--seed", type=int, default=20261017)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--ref-sample", type=int, default=0, help="queries per step of the reference arm (0 = all of them: same config as the GPU arm)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the secondary measurements (wand, maxscore, decode, pef, opt) of the default run")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else max(args.warmup, 1)
    args.queries_total = args.queries * 8          # one generated file serves every N in 1..8

    if args.impl == "reference":
        return reference_arm(args)

    # stdout carries the ONE JSON line and nothing else: libraries that print to fd 1 (NCCL's version banner under
    # NCCL_DEBUG=VERSION, for one) are pointed at stderr for the duration of the run
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(json_fd, (json.dumps(line) + "\n").encode())

    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # stdout carries the one JSON line and nothing else
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from ds2i_b200 import build
    if rank == 0:
        build.build()
    full = world == 1 and not args.no_also
    have_opt = False
    types = ["block_optpfor"]
    if full:
        types.append("block_interpolative")
        if build.builder_writes("opt"):
            types.append("opt")
            have_opt = True
    paths = ensure_data(args, rank, tuple(types))
    barrier()
    import ds2i_b200 as d
    from ds2i_b200.parallel import balanced_shards, gather_fused, query_costs, split_fused

    t0 = time.time()
    index = d.Index(paths["index"], "block_optpfor", local_rank)
    wdata = d.WandData(paths["wand"], local_rank)
    peak, peak_kind = measured_peak()

    def shard_for(total):
        """This rank's cost-balanced share of the first `total` queries of the file (+ the shard sizes of all ranks)."""
        emu = os.environ.get("DS2I_BENCH_EMULATE_SHARD")                   # diagnostic, 1 GPU: "r/W" = run the shard rank r of a W-GPU weak run has
        r, w = (int(x) for x in emu.split("/")) if emu and world == 1 else (rank, world)
        allq = d.read_queries(paths["queries"], total * (w if emu and world == 1 else 1))
        shards = balanced_shards(query_costs(index, allq, "conjunctive" if args.op in ("and", "ranked_and") else "postings"), w)
        if w != world:
            return [allq[i] for i in shards[r]], [len(shards[r])], np.asarray(shards[r])
        return [allq[i] for i in shards[rank]], [len(s) for s in shards], np.asarray(shards[rank])

    log("[bench] rank %d: index in HBM (%.1f MB) in %.1f s" % (rank, index.device_bytes() / 1e6, time.time() - t0))

    def measure(op, queries, shard_sizes, idx=index):
        """W warm-up + K timed steps of `op` over the resident batch, then the end-to-end leg."""
        batch = d.QueryBatch(idx, wdata, queries)
        pad = max(shard_sizes) * (8 + 8 * args.k) if op in d.RANKED else max(shard_sizes) * 8
        # ONE collective per step, in line with the kernels: the fused result buffer is all-gathered right behind the launch, on
        # NCCL's stream, and the next step's kernels wait for it (measured at 4 GPUs: 0.03 ms per 15 ms step).  Overlapping the
        # gather of step t with the kernels of step t + 1 (two send buffers, async_op=True; DS2I_BENCH_OVERLAP_GATHER=1) was built
        # and measured too and is SLOWER (+0.5 ms per step): the persistent query kernels fill every SM, so NCCL's kernels only
        # get scheduled in the tail of the next step anyway, and the ranks end up coupled more tightly, not less.
        overlap = bool(os.environ.get("DS2I_BENCH_OVERLAP_GATHER"))
        send = [torch.empty((pad,), dtype=torch.uint8, device="cuda") for _ in range(2)] if world > 1 else None
        gathered = [torch.empty((world, pad), dtype=torch.uint8, device="cuda") for _ in range(2)] if world > 1 else None
        pending = [None, None]
        nstep = [0]

        def step():
            # the timed launches run the kernel instances WITHOUT the algorithmic-work counters (DS2I_RUN_NO_STATS: they cost
            # registers in register-bound kernels); one more launch of the same batch, below, collects them
            if world == 1:
                return batch.run(op, args.k, stats=False)     # synchronises; returns the CUDA-event kernel time
            batch.run(op, args.k, wait=False, stats=False)    # no host synchronisation inside a step
            j = nstep[0] & 1
            nstep[0] += 1
            if not overlap:
                dist.all_gather_into_tensor(gathered[j].view(-1), batch.device_fused(pad_to=pad))      # the ONE collective of the step
                return None
            if pending[j] is not None:
                pending[j].wait()                             # stream-level: the gather that last used this buffer pair (two steps ago)
            send[j].copy_(batch.device_fused(pad_to=pad), non_blocking=True)
            pending[j] = dist.all_gather_into_tensor(gathered[j].view(-1), send[j], async_op=True)
            return None

        def drain():
            for h in pending:
                if h is not None:
                    h.wait()

        for _ in range(args.warmup):
            step()
        drain()
        barrier()
        launches0 = batch.stats()["launches"]
        kernel_ms = []
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local_rank) as clocks:
            barrier()
            ev0.record()
            for _ in range(args.steps):
                kernel_ms.append(step())
            drain()                                            # every gather of the timed steps has completed before the clock stops
            ev1.record()
            barrier()
        total_ms = ev0.elapsed_time(ev1)
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        launches = batch.stats()["launches"] - launches0
        if world > 1:                                          # kernel time of two more, synchronous steps (outside the timed region)
            kernel_ms = [batch.run(op, args.k, stats=False), batch.run(op, args.k, stats=False)]
        batch.run(op, args.k)                                  # instrumented launch of the same batch: the counters of SURVEY 8d
        stats = batch.stats()
        counts, scores = batch.fetch()
        rank_kernel_ms = [sum(kernel_ms) / len(kernel_ms)]
        if world > 1:                                          # every rank's kernel time: the step waits for the slowest shard
            tk = torch.tensor(rank_kernel_ms, dtype=torch.float64, device="cuda")
            allk = torch.empty((world,), dtype=torch.float64, device="cuda")
            dist.all_gather_into_tensor(allk, tk)
            rank_kernel_ms = [float(x) for x in allk.cpu()]
        gathered_ok = None
        if world > 1:
            # the gathered buffer of the last step holds every shard's rows: check this rank's own slice
            mine = gathered[(nstep[0] - 1) & 1][rank].cpu().numpy()
            gc, gs, _ = split_fused(mine, len(queries), args.k) if op in d.RANKED else (mine[:len(queries) * 8].view(np.uint64), None, None)
            gathered_ok = bool(np.array_equal(gc, counts) and (gs is None or np.allclose(gs, scores, rtol=1e-5, atol=0)))

        # end to end through the public call with host buffers (H2D + D2H + host-side preparation inside)
        host_buffers = d.flatten_queries(queries)
        e2e_steps = max(2, args.steps)                          # K timed calls after min(W, 3) untimed ones (the first call after the
        for _ in range(max(1, min(args.warmup, 3))):           # resident-batch steps re-grows the staging arena and wakes the host pool)
            d.query_batch(idx, wdata, op, host_buffers, args.k)
        barrier()
        te = time.perf_counter()
        for _ in range(e2e_steps):
            c2, s2, _ = d.query_batch(idx, wdata, op, host_buffers, args.k)
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - te) / e2e_steps
        te_t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te_t, op=dist.ReduceOp.MAX)
        e2e_s = float(te_t.item())
        # same results through both entry points (the union kernel's last score bit depends on the pruning history)
        assert np.array_equal(c2, counts)
        if op in d.RANKED:
            assert np.all(np.abs(s2.astype(np.float64) - scores) <= 1e-5 * np.maximum(np.abs(scores), 1e-30))
        nterms = sum(len(q) for q in queries)
        batch.close()
        return {"total_ms": total_ms, "kernel_ms": kernel_ms, "stats": stats, "launches": launches, "counts": counts, "scores": scores,
                "clocks": clocks.summary(), "e2e_s": e2e_s, "nq": len(queries), "gathered_ok": gathered_ok, "rank_kernel_ms": rank_kernel_ms,
                "h2d": nterms * 4 + (len(queries) + 1) * 8, "d2h": len(queries) * 8 + (len(queries) * args.k * 4 if op in d.RANKED else 0)}

    def roofline_of(m, op):
        st = m["stats"]
        # algorithmic bytes: compressed payload of the decoded blocks + 8 B of block metadata (max +
        # endpoint) per decoded docs block + one 4-B norm_len per scored document.  The kernel's own block_max
        # probes (32 entries per ballot step) are implementation traffic and are NOT counted.
        alg = st["docs_bytes"] + st["freqs_bytes"] + 8 * st["docs_blocks"] + 4 * st["docs_scored"]
        kern = sum(m["kernel_ms"]) / len(m["kernel_ms"])
        ach = alg / (kern * 1e-3) / 1e9
        return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic_per_launch(op),
                "traffic_kind": "static: ncu --set full capture committed under profiles/, not measured in this run",
                "counters_from": "one more launch of the same batch with the instrumented kernel instance (the timed launches run without counters)",
                "peak_kind": peak_kind, "algorithmic_bytes_per_launch": alg, "kernel_ms": kern, "counters": st}

    total_q = args.queries * world if args.scaling == "weak" else args.queries
    queries, shard_sizes, qidx = shard_for(total_q)
    m = measure(args.op, queries, shard_sizes)
    also_m = {}
    if args.op == "ranked_and" and not args.no_also:
        for op2 in ("wand", "maxscore") + (("or", "ranked_or") if world == 1 else ()):
            also_m[op2] = measure(op2, queries, shard_sizes)            # the second half of BASELINE.json's metric, same protocol
    strong_m = None
    if world > 1 and args.scaling == "weak" and not args.no_also:
        sq, ssz, _ = shard_for(args.queries)                               # strong scaling: the 10k-query batch of N = 1 cut over the ranks
        strong_m = {op2: measure(op2, sq, ssz) for op2 in ((args.op, "wand") if args.op == "ranked_and" else (args.op,))}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    def line_of(mm, op, nq_total, scaling, itype="block_optpfor"):
        ms_per_step = mm["total_ms"] / args.steps
        return {"metric": METRIC % (op, itype), "value": nq_total / (ms_per_step * 1e-3), "unit": "queries/s", "ms_per_step": ms_per_step,
                "scaling": scaling, "e2e": {"value": nq_total / mm["e2e_s"], "unit": "queries/s", "h2d_bytes_per_step": mm["h2d"], "d2h_bytes_per_step": mm["d2h"]},
                "gpu_launches": mm["launches"], "roofline": roofline_of(mm, op), "clocks": mm["clocks"], "kernel_ms_per_rank": mm["rank_kernel_ms"]}

    nq_total = sum(shard_sizes)
    head = line_of(m, args.op, nq_total, args.scaling)
    line = {
        "metric": head["metric"], "value": head["value"], "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "u32+f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "op": args.op, "index_type": "block_optpfor",
                   "num_docs": args.docs, "num_terms": args.terms, "queries_per_step": nq_total, "queries_per_gpu_per_step": len(queries), "k": args.k,
                   "seed": args.seed, "index_bytes": index.device_bytes(),
                   "l2": "no flush: index (%.0f MB) and per-step touched bytes exceed the 126 MB L2" % (index.device_bytes() / 1e6),
                   "parallelism": "queries dealt cost-balanced over %d GPU(s), index replicated, ONE NCCL all_gather of the fused per-shard results per step, "
                                  "enqueued behind the kernels without a host synchronisation" % world},
        "clocks": head["clocks"], "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "roofline": head["roofline"],
        "kernel_ms_per_rank": head["kernel_ms_per_rank"],
    }
    if m["gathered_ok"] is not None:
        line["gathered_results_ok"] = m["gathered_ok"]
    for op2, m2 in also_m.items():
        line.setdefault("also", {})[op2] = line_of(m2, op2, nq_total, args.scaling)
    if strong_m:
        line.setdefault("also", {})["strong"] = {op2: dict(line_of(m2, op2, args.queries, "strong"), queries_per_gpu=m2["nq"],
                                                            note="the 10k-query batch of N = 1 dealt over the %d GPUs; the limiter is the tail: the costliest "
                                                                 "queries' work items and the fixed per-step launch + gather latency" % world)
                                                 for op2, m2 in strong_m.items()}

    if not args.no_cpu_baseline and world == 1:
        nq = len(queries)
        try:
            line["cpu_baseline"] = cpu_baseline_for(paths, args.op, nq)
            line["parity"] = parity_block(d, paths, args.op, args.k, qidx, m["counts"], m["scores"])
            for op2, m2 in also_m.items():
                line["also"][op2]["cpu_baseline"] = cpu_baseline_for(paths, op2, nq, single_prefix=200 if op2 in ("or", "ranked_or") else 1000)
                line["also"][op2]["parity"] = parity_block(d, paths, op2, args.k, qidx, m2["counts"], m2["scores"])
            # the same roofline with the bytes the REFERENCE algorithm decodes (its own block_profiler) instead of the
            # device counters: what SURVEY.md 8d defines as algorithmic bytes
            for op2, tgt in [(args.op, line)] + [(o, line["also"][o]) for o in also_m]:
                try:
                    prof = reference_block_profile(paths, op2, 500 if op2 in ("or", "ranked_or") else 2000)
                    kern = tgt["roofline"]["kernel_ms"]
                    prof["achieved_GBps"] = prof["bytes_per_query"] * nq / (kern * 1e-3) / 1e9
                    prof["frac"] = prof["achieved_GBps"] / peak
                    prof["sample"] = "block_profiler over the first %d queries, scaled to the %d of the step" % (prof["queries"], nq)
                    tgt["roofline"]["reference_block_profiler"] = prof
                except Exception as e:
                    tgt["roofline"]["reference_block_profiler"] = {"failed": repr(e)}
        except Exception as e:   # the baseline is a report, never a reason to lose the measurement
            line["cpu_baseline"] = {"value": None, "unit": "queries/s", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %r" % (e,)}

    if full:
        index.close()
        for itype in ("block_optpfor", "block_interpolative"):
            try:
                line.setdefault("also", {})["decode_" + itype] = decode_leg(d, args, paths, itype, peak, args.steps, args.warmup)
            except Exception as e:
                line.setdefault("also", {})["decode_" + itype] = {"failed": repr(e)}
        if have_opt:
            try:
                from ds2i_b200 import bench_opt
                line["also"].update(bench_opt.legs(d, args, paths, queries, qidx, wdata, peak, measure, line_of, cpu_baseline_for, parity_block, ref_tool))
            except Exception as e:
                line["also"]["pef"] = {"failed": repr(e)}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
