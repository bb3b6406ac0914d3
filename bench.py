#!/usr/bin/env python3
"""Benchmark of the hot path: one step = one pass of a query operator over a 10k-query batch on the
synthetic 10M-doc / 1M-term Zipfian index (BASELINE.json configs[3]/[4]), index resident in HBM.

  python bench.py --gpus N --steps K --warmup W [--op ranked_and|wand|maxscore|and]
  python bench.py --impl reference ...     times the compiled reference (oracle/_ref) on the host cores

value   queries/s with the batch already resident in HBM (CUDA events around the K steps, max over ranks)
e2e     queries/s through ds2i_gpu_query_batch with HOST buffers (H2D of the queries and D2H of
        counts + top-k inside the timed region, host-side query preparation included)
roofline  algorithmic bytes (device counters = blocks/bytes the reference algorithm decodes)
          / average kernel time, against MEASURED_PEAKS.json hbm_gbs
cpu_baseline  the reference's own operators (oracle/_ref/ref_tool, compiled from /root/reference) on all
          host cores over a bounded sample of the same queries, same index file.
Multi-GPU: index replicated per GPU, every rank evaluates its own 10k-query shard (weak scaling), the
per-shard top-k is gathered to every rank with one NCCL all_gather inside the timed region.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "queries/sec (%s, block_optpfor, synthetic 10M-doc/1M-term Zipfian index, 10k-query batch, top-10)"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def data_dir(args):
    base = os.environ.get("DS2I_BENCH_DATA", "/tmp/ds2i_b200_data")
    return os.path.join(base, "S_%d_%d_%d_q%d" % (args.docs, args.terms, args.seed, args.queries_total))


def ensure_data(args, rank=0):
    """Synthetic collection -> ds2i-format index + wand data + queries, built by OUR builder
    (ds2i_b200/csrc/builder.cpp; byte-identical to the reference's create_freq_index output)."""
    d = data_dir(args)
    done = os.path.join(d, "DONE")
    if not os.path.exists(done):
        if rank == 0:
            from ds2i_b200 import build
            build.build()
            os.makedirs(d, exist_ok=True)
            t0 = time.time()
            subprocess.run([build.BUILDER, "synth", os.path.join(d, "S"), str(args.docs), str(args.terms), str(args.seed), "0",
                            str(args.queries_total)], check=True)
            log("[bench] built synthetic index in %.1f s -> %s" % (time.time() - t0, d))
            open(done, "w").write("ok\n")
        else:
            while not os.path.exists(done):
                time.sleep(1.0)
    return {"index": os.path.join(d, "S.block_optpfor.idx"), "wand": os.path.join(d, "S.wand"), "queries": os.path.join(d, "S.queries")}


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
        except Exception as e:          # NVML missing: fall back to one nvidia-smi query
            try:
                r = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"],
                                   stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5)
                a, b = [float(x) for x in r.stdout.strip().split(",")]
                self.samples.append((a, 0)); self.max_mhz = b
            except Exception:
                pass

    def __enter__(self):
        self.t.start()
        time.sleep(0.05)
        return self

    def __exit__(self, *a):
        self.stop = True
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
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        return json.load(open(p)).get(op)
    return None


def run_reference_tool(paths, op, threads, sample, passes):
    tool = os.path.join(ROOT, "oracle", "_ref", "ref_tool")
    if not os.access(tool, os.X_OK):
        raise RuntimeError("oracle/_ref/ref_tool missing (built in the build container by oracle/Makefile)")
    r = subprocess.run([tool, "bench", "block_optpfor", paths["index"], paths["wand"], paths["queries"], op, str(threads), str(sample), str(passes)],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, check=True)
    return json.loads(r.stdout.strip().splitlines()[-1])


def reference_block_profile(paths, op, sample):
    """Bytes the REFERENCE algorithm decodes per query: ds2i's own block_profiler (block_profiler.hpp:40-54,
    what profile_queries.cpp uses) run over the first `sample` queries by oracle/_ref/ref_tool.  B_q = docs payload of the
    blocks decoded at least once + freqs payload likewise + 8 B (block_max + endpoint) per decoded docs block."""
    import numpy as np
    tool = os.path.join(ROOT, "oracle", "_ref", "ref_tool")
    out = os.path.join(os.path.dirname(paths["index"]), "profile.%s.bin" % op)
    subprocess.run([tool, "profile", "block_optpfor", paths["index"], paths["wand"], paths["queries"], op, out, str(sample)],
                   stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, check=True, timeout=300)
    raw = np.fromfile(out, dtype="<u8")
    nq = int(raw[0])
    a = raw[1:1 + 6 * nq].reshape(nq, 6).astype(np.float64)
    return {"queries": nq, "docs_blocks_per_query": float(a[:, 0].mean()), "freqs_blocks_per_query": float(a[:, 1].mean()),
            "bytes_per_query": float((a[:, 2] + a[:, 3] + 8 * a[:, 0]).mean()), "list_bytes_per_query": float(a[:, 4].mean())}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    paths = ensure_data(args)
    cores = os.cpu_count() or 1
    sample = min(args.ref_sample, args.queries)
    out = run_reference_tool(paths, args.op, cores, sample, args.warmup + args.steps)
    secs = out["pass_seconds"][args.warmup:]
    per_step = sum(secs) / len(secs)
    qps = sample / per_step
    line = {
        "impl": "reference", "metric": METRIC % args.op, "value": qps, "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32+f32", "data": "synthetic",
        "config": {"workload": "ranked top-10 over synthetic Zipfian index (configs[3]/[4])", "op": args.op, "index_type": "block_optpfor",
                   "num_docs": args.docs, "num_terms": args.terms, "queries_per_step": sample, "seed": args.seed},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "reference",
                         "sample": "first %d of the %d queries per step, ds2i %s_query compiled from /root/reference (-O3 -march=x86-64-v3), thread t takes queries t,t+n,.. (profile_queries.cpp:21-39)" % (sample, args.queries, args.op)},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--op", default="ranked_and")
    ap.add_argument("--docs", type=int, default=10_000_000)
    ap.add_argument("--terms", type=int, default=1_000_000)
    ap.add_argument("--queries", type=int, default=10_000, help="queries per GPU per step")
    ap.add_argument("--seed", type=int, default=20261017)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--ref-sample", type=int, default=2000, help="queries per step of the CPU reference runs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the secondary wand measurement of the default run")
    ap.add_argument("--check", type=int, default=200, help="queries checked against the reference in the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else max(args.warmup, 1)
    args.queries_total = args.queries * 8          # one generated file serves every N in 1..8 (rank r takes slice r)

    if args.impl == "reference":
        return reference_arm(args)

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
    paths = ensure_data(args, rank)
    barrier()
    import ds2i_b200 as d
    from ds2i_b200.parallel import shard_queries, gather_topk

    t0 = time.time()
    index = d.Index(paths["index"], "block_optpfor", local_rank)
    wdata = d.WandData(paths["wand"], local_rank)
    all_queries = d.read_queries(paths["queries"], args.queries * world)
    queries = shard_queries(all_queries, rank, world, args.queries)
    log("[bench] rank %d: index in HBM (%.1f MB) + %d queries in %.1f s" % (rank, index.device_bytes() / 1e6, len(queries), time.time() - t0))
    batch = d.QueryBatch(index, wdata, queries)

    host_buffers = d.flatten_queries(queries)          # the caller's host buffers (terms, offsets)
    nterms = sum(len(q) for q in queries)
    h2d = nterms * 4 + (len(queries) + 1) * 8
    d2h = len(queries) * 8 + len(queries) * args.k * 4

    def measure(op):
        """W warm-up + K timed steps of `op` over the resident batch, then the end-to-end leg."""
        def step():
            ms = batch.run(op, args.k)
            if world > 1:
                counts_t, scores_t = batch.device_results(args.k)
                gather_topk(counts_t, scores_t, world)
            return ms

        for _ in range(args.warmup):
            step()
        launches0 = batch.stats()["launches"]
        kernel_ms = []
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local_rank) as clocks:
            barrier()
            ev0.record()
            for _ in range(args.steps):
                kernel_ms.append(step())
            ev1.record()
            barrier()
        total_ms = ev0.elapsed_time(ev1)
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        stats = batch.stats()
        counts, scores = batch.fetch()

        # end to end through the public call with host buffers (H2D + D2H + host-side preparation inside)
        e2e_steps = max(2, min(args.steps, 3))
        d.query_batch(index, wdata, op, host_buffers, args.k)
        barrier()
        te = time.perf_counter()
        for _ in range(e2e_steps):
            c2, s2, _ = d.query_batch(index, wdata, op, host_buffers, args.k)
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - te) / e2e_steps
        te_t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te_t, op=dist.ReduceOp.MAX)
        e2e_s = float(te_t.item())
        # same results through both entry points (the union kernel's last score bit depends on the pruning history)
        assert np.array_equal(c2, counts)
        assert np.all(np.abs(s2.astype(np.float64) - scores) <= 1e-5 * np.maximum(np.abs(scores), 1e-30))
        return {"total_ms": total_ms, "kernel_ms": kernel_ms, "stats": stats, "launches": stats["launches"] - launches0,
                "counts": counts, "scores": scores, "clocks": clocks.summary(), "e2e_s": e2e_s}

    m = measure(args.op)
    also = {}
    if args.op == "ranked_and" and not args.no_also:
        also["wand"] = measure("wand")            # the second half of BASELINE.json's metric, same protocol
    total_ms, kernel_ms, stats, counts, scores, clocks_summary, e2e_s = (m["total_ms"], m["kernel_ms"], m["stats"], m["counts"],
                                                                          m["scores"], m["clocks"], m["e2e_s"])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_per_step = total_ms / args.steps
    nq_total = len(queries) * world
    value = nq_total / (ms_per_step * 1e-3)
    # algorithmic bytes: compressed payload of the decoded blocks + 8 B of block metadata (max +
    # endpoint) per decoded docs block + one 4-B norm_len per scored document.  The kernel's own block_max
    # probes (32 entries per ballot step) are implementation traffic and are NOT counted.
    alg_bytes = stats["docs_bytes"] + stats["freqs_bytes"] + 8 * stats["docs_blocks"] + 4 * stats["docs_scored"]
    kern_ms = sum(kernel_ms) / len(kernel_ms)
    peak, peak_kind = measured_peak()
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC % args.op, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32+f32",
        "data": "synthetic",
        "config": {"workload": "ranked top-10 over synthetic Zipfian index (configs[3]/[4])", "op": args.op, "index_type": "block_optpfor",
                   "num_docs": args.docs, "num_terms": args.terms, "queries_per_gpu_per_step": len(queries), "k": args.k,
                   "seed": args.seed, "index_bytes": index.device_bytes(),
                   "l2": "no flush: index (%.0f MB) and per-step touched bytes exceed the 126 MB L2" % (index.device_bytes() / 1e6),
                   "parallelism": "query batch sharded over %d GPU(s), index replicated, NCCL all_gather of top-k" % world},
        "clocks": clocks_summary,
        "e2e": {"value": nq_total / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": m["launches"],
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic_per_launch(args.op),
                     "peak_kind": peak_kind, "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kern_ms,
                     "counters": stats},
    }

    for op2, m2 in also.items():
        ms2 = m2["total_ms"] / args.steps
        line.setdefault("also", {})[op2] = {"metric": METRIC % op2, "value": nq_total / (ms2 * 1e-3), "unit": "queries/s", "ms_per_step": ms2,
                                            "e2e": {"value": nq_total / m2["e2e_s"], "unit": "queries/s"}, "gpu_launches": m2["launches"],
                                            "counters": m2["stats"]}

    if not args.no_cpu_baseline and world == 1:
        try:
            cores = os.cpu_count() or 1
            sample = min(args.ref_sample, len(queries))
            out = run_reference_tool(paths, args.op, cores, sample, 2)
            line["cpu_baseline"] = {"value": sample / out["pass_seconds"][-1], "unit": "queries/s", "cores": cores, "kind": "reference",
                                    "sample": "first %d queries, second of 2 passes, ds2i %s_query compiled from the reference sources, all host threads" % (sample, args.op)}
            for op2 in also:
                out2 = run_reference_tool(paths, op2, cores, sample, 2)
                line["also"][op2]["cpu_baseline"] = {"value": sample / out2["pass_seconds"][-1], "unit": "queries/s", "cores": cores, "kind": "reference",
                                                     "sample": "first %d queries, second of 2 passes, all host threads" % sample}
            # the same roofline with the bytes the REFERENCE algorithm decodes (its own block_profiler) instead of the
            # device counters: what SURVEY.md 8d defines as algorithmic bytes
            try:
                prof = reference_block_profile(paths, args.op, sample)
                ref_bytes = prof["bytes_per_query"] * len(queries)
                prof["achieved_GBps"] = ref_bytes / (kern_ms * 1e-3) / 1e9
                prof["frac"] = prof["achieved_GBps"] / peak
                prof["device_counted_bytes_per_query"] = alg_bytes / len(queries)
                prof["sample"] = "block_profiler over the first %d queries, scaled to the %d of the step" % (prof["queries"], len(queries))
                line["roofline"]["reference_block_profiler"] = prof
            except Exception as e:
                line["roofline"]["reference_block_profiler"] = {"failed": repr(e)}
            # parity at full size: the reference's own results for the first queries (the checker, not the product)
            ncheck = min(args.check, len(queries))
            if ncheck:
                tool = os.path.join(ROOT, "oracle", "_ref", "ref_tool_strict")
                tmp = os.path.join(data_dir(args), "check.%s.bin" % args.op)
                subprocess.run([tool, "dump", "block_optpfor", paths["index"], paths["wand"], paths["queries"], tmp, args.op, str(args.k), str(ncheck)], check=True)
                sys.path.insert(0, os.path.join(ROOT, "tests"))
                from util import load_dump
                ec, es = load_dump(tmp, (args.op,))[args.op]
                ok_counts = bool(np.array_equal(ec, counts[:ncheck]))
                ok_scores = bool(np.array_equal(es.view(np.uint32), scores[:ncheck].view(np.uint32)))
                rel = float(np.max(np.abs(es.astype(np.float64) - scores[:ncheck]) / np.maximum(np.abs(es), 1e-30)))
                line["parity"] = {"queries_checked": ncheck, "counts_bit_exact": ok_counts, "scores_bit_exact": ok_scores,
                                  "scores_max_rel_err": rel, "tolerance": 1e-5, "against": "ds2i reference, -ffp-contract=off build"}
        except Exception as e:   # the baseline is a report, never a reason to lose the measurement
            line["cpu_baseline"] = {"value": None, "unit": "queries/s", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %r" % (e,)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
