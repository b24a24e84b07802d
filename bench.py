#!/usr/bin/env python
"""bench.py -- hybrid (dense + BM25 + MMR + RRF k=60) top-10 retrieval over a
synthetic 10M x 768 corpus on 1..8 B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # one JSON line (rank 0)
  python bench.py --impl reference --steps K --warmup W    # CPU port of the reference path

A step is one batch of `--batch` hybrid queries through the hot path.
`value` is whole-job QPS with queries already resident in HBM; `e2e` is the same
through the host-buffer call (pinned H2D of the queries + D2H of the results
inside the timed region, CUDA graph replay); `latency` is the single-query
(B=1) end-to-end latency distribution.  The corpus is fixed, so adding GPUs is
strong scaling: rows are sharded, only per-shard top-k lists are exchanged.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "hybrid_top10_qps"
UNIT = "queries/s"
VOCAB = 30000
MEAN_LEN = 64
TOP_K = 10


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--cpu-sample-rows", type=int, default=20000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--latency-iters", type=int, default=1000)
    ap.add_argument("--no-c2", action="store_true", help="skip the C2 side measurement (1M x 768 dense, B=1 / B=1024)")
    return ap.parse_args()


def workload_name(a):
    return (f"{a.rows / 1e6:g}M x {a.dim} hybrid (dense bf16 exact top-k + MMR pool 24 + BM25 V={VOCAB} "
            f"mean_len={MEAN_LEN} + RRF k=60) top-{TOP_K}, batch {a.batch} queries/step")


# --------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md: sample during the timed region)
# --------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self, settle_s: float = 0.6):
        """nvidia-smi needs a few hundred ms before its first sample: start early, then
        mark() the beginning of the region whose samples count."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            time.sleep(settle_s)
        except Exception:
            self.proc = None
        self.first = 0

    def mark(self):
        self.first = len(self.rows)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows[self.first:]:
            try:
                sm.append(float(r[0]))
                smax.append(float(r[1]))
                for name, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------
# reference arm: the CPU port of the reference path on a bounded sample
# --------------------------------------------------------------------------
def _cpu_sample(a, n_sample):
    import torch
    from classmate_rag_b200 import synth
    from oracle.cpu_baseline import ReferencePort
    n_sample = min(n_sample, a.rows)
    emb = synth.dense_corpus(a.rows, a.dim, "cpu", row_lo=0, row_hi=n_sample).float().numpy()
    doc_ptr, tokens = synth.lexical_corpus(a.rows, VOCAB, MEAN_LEN, "cpu", doc_lo=0, doc_hi=n_sample)
    ptr, tok = doc_ptr.numpy(), tokens.numpy()
    words = np.array([f"w{chr(97 + i % 26)}{chr(97 + (i // 26) % 26)}{chr(97 + (i // 676) % 26)}{chr(97 + i // 17576)}"
                      for i in range(VOCAB)])
    docs = [words[tok[ptr[i]:ptr[i + 1]]].tolist() for i in range(n_sample)]
    port = ReferencePort(emb, docs)
    nq = 64
    q = torch.nn.functional.normalize(torch.from_numpy(emb[:nq]) + 0.5 * torch.randn(nq, a.dim) / a.dim ** 0.5, dim=1).numpy()
    texts = [" ".join(words[[t for t in terms if t >= 0]]) for terms in synth.lexical_queries(nq, VOCAB)]
    return port, q, texts, n_sample


def cpu_baseline(a, budget_s=20.0):
    from oracle.cpu_baseline import host_cores, time_reference_port
    port, q, texts, n_sample = _cpu_sample(a, a.cpu_sample_rows)
    sec, n = time_reference_port(port, q, texts, TOP_K, budget_s=budget_s)
    qps_sample = 1.0 / sec
    return {"value": qps_sample * n_sample / a.rows, "unit": UNIT, "cores": host_cores(), "kind": "port",
            "sample": (f"first {n_sample} rows/docs of the workload, {n} single queries, median {sec * 1e3:.1f} ms/query "
                       f"({qps_sample:.3f} QPS on the sample); the reference path is O(rows) per query (per-query BM25Okapi "
                       f"rebuild + full scan), so QPS is scaled linearly by {n_sample}/{a.rows}; dense stage is an exact "
                       f"NumPy fp32 stand-in for Chroma HNSW (chromadb not installable offline)"),
            "qps_on_sample": qps_sample}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.cpu_baseline import host_cores
    port, q, texts, n_sample = _cpu_sample(a, a.cpu_sample_rows)
    times = []
    for i in range(a.warmup + a.steps):
        t0 = time.perf_counter()
        port.retrieve(q[i % len(q)], texts[i % len(texts)], TOP_K)
        if i >= a.warmup:
            times.append(time.perf_counter() - t0)
    ms = float(np.mean(times) * 1e3)
    qps = (1e3 / ms) * n_sample / a.rows
    cb = {"value": qps, "unit": UNIT, "cores": host_cores(), "kind": "port",
          "sample": f"each step = 1 query on the first {n_sample} rows/docs; QPS scaled linearly by {n_sample}/{a.rows} (O(rows) path)"}
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
                      "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
                      "vs_baseline": None, "dtype": "f32/f64", "data": "synthetic",
                      "config": {"workload": workload_name(a)}, "cpu_baseline": cb,
                      "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# --------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------
def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
        return
    import torch
    import torch.distributed as dist
    from classmate_rag_b200 import lexical, ops, sharding, synth
    from classmate_rag_b200.engine import GraphedSearch, HybridEngine, PipelinedSearch, SearchParams

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sm, cc_major, cc_minor = ops.device_info()

    # ---- build this rank's shard -----------------------------------------
    t_build = time.perf_counter()
    lo, hi = sharding.shard_range(a.rows, rank, world)
    emb = synth.dense_corpus(a.rows, a.dim, dev, row_lo=lo, row_hi=hi)
    doc_ptr, tokens = synth.lexical_corpus(a.rows, VOCAB, MEAN_LEN, dev, doc_lo=lo, doc_hi=hi)
    if world > 1:
        tok_counts = torch.zeros(world, dtype=torch.int64, device=dev)
        tok_counts[rank] = tokens.numel()
        dist.all_reduce(tok_counts)
        stats = sharding.global_corpus_stats(doc_ptr, tokens, VOCAB, doc_lo=lo, n_docs_total=a.rows,
                                             token_offset=int(tok_counts[:rank].sum()))
    else:
        stats = lexical.corpus_stats(doc_ptr, tokens, VOCAB)
    lex = lexical.build_lexical_index(doc_ptr, tokens, VOCAB, stats=stats)
    del doc_ptr, tokens
    torch.cuda.empty_cache()
    comm = sharding.ShardComm() if world > 1 else None
    eng = HybridEngine(emb, lex, row_offset=lo, comm=comm)
    p = SearchParams(top_k=TOP_K)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t_build

    # ---- queries -----------------------------------------------------------
    n_steps = a.warmup + a.steps
    nq = a.batch * n_steps
    q_f32, planted = synth.dense_queries(a.rows, a.dim, nq, dev)
    terms = synth.lexical_queries(nq, VOCAB)
    q_bf16 = ops.f32_to_bf16(q_f32)
    dev_terms = []
    for s in range(n_steps):
        qt, qp = lexical.pack_queries(terms[s * a.batch:(s + 1) * a.batch])
        dev_terms.append((qt.to(dev), qp.to(dev)))
    posting_bytes = [sum(lex.posting_bytes(t) for t in terms[s * a.batch:(s + 1) * a.batch]) for s in range(n_steps)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value: device-resident inputs, CUDA events, max over ranks -------------
    # The step is replayed from its CUDA graph (one launch per step); the queries of every
    # step already live in HBM and are copied device-to-device into the graph's input buffers.
    clocks = ClockSampler(local)
    q_res = q_f32.contiguous()
    gv = GraphedSearch(eng, p, a.batch, max_terms=16)
    stream = gv.stream

    def resident_step(s):
        gv.launch_resident(q_res[s * a.batch:(s + 1) * a.batch], *dev_terms[s])
    dense_ms, lex_ms = [], []
    for s in range(a.warmup):
        resident_step(s)
    clocks.start()
    barrier()
    clocks.mark()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for s in range(a.warmup, n_steps):
        resident_step(s)
    ev1.record(stream)
    barrier()
    total_ms = max_over_ranks(ev0.elapsed_time(ev1))
    ms_per_step = total_ms / a.steps
    value = a.batch * a.steps / (total_ms * 1e-3)
    stream = torch.cuda.current_stream()

    # ---- per-stage device time.  Single shard: the same steps again with timing events at the
    # stage boundaries INSIDE the step (a stage timed back to back on its own runs at other
    # clocks: twenty BM25 launches in a row hit the power cap, inside a step they alternate
    # with the memory-bound dense pass).  Sharded: each retriever back to back (its own exchange).
    pool = p.pool
    overlap_on = eng.overlap
    eng.overlap = False   # stage times are those of each retriever running alone inside a (serial) step
    if world == 1:
        marks = []
        for s in range(a.warmup, n_steps):
            ev = []
            eng.search(q_bf16[s * a.batch:(s + 1) * a.batch], *dev_terms[s], p, stage_events=ev)
            marks.append(ev)
        torch.cuda.synchronize()
        dense_ms = [m[0].elapsed_time(m[1]) for m in marks]
        lex_ms = [m[2].elapsed_time(m[3]) for m in marks]
        stage_step_ms = float(np.mean([m[0].elapsed_time(m[4]) for m in marks]))   # this (eager) loop's own step
    else:
        def stage_ms(fn):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            fn(a.warmup)
            torch.cuda.synchronize()
            e0.record(stream)
            for s in range(a.warmup, n_steps):
                fn(s)
            e1.record(stream)
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / a.steps
        dense_ms = [stage_ms(lambda s: eng.dense_pool(q_bf16[s * a.batch:(s + 1) * a.batch], pool))]
        lex_ms = [stage_ms(lambda s: eng.lexical_topk(*dev_terms[s], p.k_bm25))]
        stage_step_ms = ms_per_step
    eng.overlap = overlap_on
    clock_info = clocks.stop()
    dense_avg = float(np.mean(dense_ms))
    lex_avg = float(np.mean(lex_ms))

    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        hbm_peak, peak_src = float(json.loads(peaks_path.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
    else:
        hbm_peak, peak_src = 6650.0, "B200_PROFILING.md fallback 6.65 TB/s"
    # the matrix is read once per 128 queries (tcgen05 path, batch > 8) or per 8 (scan path)
    dense_bytes = (hi - lo) * a.dim * 2
    passes = (a.batch + 127) // 128 if a.batch > 8 else 1
    achieved = dense_bytes * passes / (dense_avg * 1e-3) / 1e9
    dense_kernel = ("dense_mma_kernel<MAIN> (tcgen05/TMA; events bracket cmr_dense_topk = sample pass over every 16th/32nd "
                    "tile + bound + main pass + finalize)" if a.batch > 8 else
                    "dense_scan_kernel (events bracket cmr_dense_topk = scan + finalize)")
    traffic = None
    tpath = ROOT / "profiles" / "traffic.json"
    if tpath.exists():
        try:
            t = json.loads(tpath.read_text())
            key = f"{'dense_mma_main' if a.batch > 8 else 'dense_scan'}:{hi - lo}x{a.dim}"
            traffic = t.get(key)
        except Exception:
            traffic = None
    lex_bytes = float(np.mean(posting_bytes[a.warmup:]))
    roofline = {"bound": "hbm", "kernel": dense_kernel,
                "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": dense_bytes,
                "avg_launch_ms": dense_avg, "share_of_step": dense_avg / stage_step_ms,
                "bm25": {"kernel": "bm25_tile_kernel (+finalize)", "algorithmic_bytes_per_step": lex_bytes,
                         "avg_ms_per_step": lex_avg, "achieved": lex_bytes / (lex_avg * 1e-3) / 1e9, "unit": "GB/s",
                         "frac": lex_bytes / (lex_avg * 1e-3) / 1e9 / hbm_peak, "share_of_step": lex_avg / stage_step_ms}}

    # ---- e2e: host buffers in, host results out, every step ---------------------
    q_host = q_f32.cpu().numpy()
    gs = PipelinedSearch(eng, p, a.batch, max_terms=16)
    for s in range(a.warmup):
        gs.submit(q_host[s * a.batch:(s + 1) * a.batch], terms[s * a.batch:(s + 1) * a.batch])
    gs.drain()
    barrier()
    t0 = time.perf_counter()
    last = None
    for s in range(a.warmup, n_steps):
        # every step: pinned host inputs -> H2D -> graph replay -> D2H; the host stages step s+1
        # while the device runs step s (results of step s-1 are handed back by this call)
        got = gs.submit(q_host[s * a.batch:(s + 1) * a.batch], terms[s * a.batch:(s + 1) * a.batch])
        last = got if got is not None else last
    last = gs.drain()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e = {"value": a.batch * a.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": gs.h2d_bytes,
           "d2h_bytes_per_step": gs.d2h_bytes, "ms_per_step": e2e_s / a.steps * 1e3,
           "path": "PipelinedSearch: pinned host queries -> H2D -> CUDA-graph replay of the kernel sequence -> D2H "
                   "results, every step; two graph slots so the host stages step s+1 while the device runs step s"}
    # sanity on the last batch: the planted row is the dense top-1 unless MMR/RRF reorder it out of the top-10
    ids_last = last[0]
    planted_last = planted[(n_steps - 1) * a.batch:].numpy()
    hit = float(np.mean([planted_last[i] in ids_last[i] for i in range(a.batch)]))

    # ---- single-query latency (B=1), end to end: host clock and CUDA events -----------------
    g1 = GraphedSearch(eng, p, 1, max_terms=16)
    lat, lat_dev = [], []
    for i in range(min(50, max(nq, 1) * 4)):
        g1(q_host[i % nq:i % nq + 1], [terms[i % nq]])
    barrier()
    for i in range(a.latency_iters):
        j = i % nq
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t1 = time.perf_counter()
        g1.set_queries(q_host[j:j + 1], [terms[j]])
        d0.record(g1.stream)
        g1.launch()
        d1.record(g1.stream)
        g1.stream.synchronize()
        lat.append((time.perf_counter() - t1) * 1e3)
        lat_dev.append(d0.elapsed_time(d1))
    lat, lat_dev = np.array(lat), np.array(lat_dev)
    latency = {"batch": 1, "iters": int(a.latency_iters), "p50_ms": float(np.percentile(lat, 50)),
               "p95_ms": float(np.percentile(lat, 95)), "p99_ms": float(np.percentile(lat, 99)),
               "device_p50_ms": float(np.percentile(lat_dev, 50)), "device_p99_ms": float(np.percentile(lat_dev, 99)),
               "qps_serial": float(1e3 / np.mean(lat)),
               "hbm_frac_at_p50": (hi - lo) * a.dim * 2 / (float(np.percentile(lat, 50)) * 1e-3) / 1e9 / hbm_peak}

    # ---- side measurement, BASELINE config C2: 1M x 768 exact dense top-10, batch 1 and 1024 ----
    c2 = None
    if world == 1 and not a.no_c2:
        peaks = json.loads(peaks_path.read_text()) if peaks_path.exists() else {}
        tf_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        n2, b2 = 1_000_000, 1024
        emb2 = synth.dense_corpus(n2, a.dim, dev)
        q2 = ops.f32_to_bf16(synth.dense_queries(n2, a.dim, b2, dev)[0])

        def timed(qb, algo, iters=20):
            ws = ops.DenseWorkspace(n2, a.dim, qb.shape[0], TOP_K, dev)
            for _ in range(3):
                ops.dense_topk(emb2, qb, TOP_K, workspace=ws, algo=algo)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
            ev[0].record(stream)
            for i in range(iters):
                ops.dense_topk(emb2, qb, TOP_K, workspace=ws, algo=algo)
                ev[i + 1].record(stream)
            torch.cuda.synchronize()
            return float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(iters)]))
        t_b1 = timed(q2[:1], "scan")
        t_b1024 = timed(q2, "mma")
        flops = 2.0 * b2 * n2 * a.dim
        c2 = {"workload": f"1M x {a.dim} exact dense top-{TOP_K}",
              "batch1": {"kernel": "dense_scan_kernel", "ms": t_b1, "qps": 1e3 / t_b1,
                         "hbm_gbs": n2 * a.dim * 2 / (t_b1 * 1e-3) / 1e9, "hbm_frac": n2 * a.dim * 2 / (t_b1 * 1e-3) / 1e9 / hbm_peak},
              "batch1024": {"kernel": "dense_mma_kernel (tcgen05/TMA; sample + bound + main + finalize)", "ms": t_b1024,
                            "qps": b2 * 1e3 / t_b1024, "tflops": flops / (t_b1024 * 1e-3) / 1e12,
                            "tensor_frac": flops / (t_b1024 * 1e-3) / 1e12 / tf_peak,
                            "peak_tflops": tf_peak, "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained"
                            if peaks else "B200_PROFILING.md fallback (sustained)"}}
        del emb2, q2

    # kernels per step (resident loop): dense = sample pass, bound, main pass, finalize (tcgen05 path)
    # or scan, finalize; then gather, mmr, bm25 tile, bm25 finalize, fuse; sharded: + 2 merges
    launches_per_step = 1 + (4 if a.batch > 8 else 2) + 5 + (2 if world > 1 else 0)   # +1: f32 -> bf16 of the queries
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
           "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "bf16", "data": "synthetic",
           "config": {"workload": workload_name(a), "rows": a.rows, "dim": a.dim, "batch": a.batch, "top_k": TOP_K,
                      "vocab": VOCAB, "postings_this_rank": lex.n_postings, "rows_this_rank": hi - lo,
                      "parallelism": f"row-shard x{world}", "sm_count": sm, "cc": f"{cc_major}.{cc_minor}",
                      "l2": "inputs larger than L2 (matrix %.1f GB per rank)" % (dense_bytes / 1e9),
                      "index_build_s": build_s,
                      "overlap": ("BM25 kernels on a side stream beside the dense scan inside the step's CUDA graph"
                                  if overlap_on else "off (CMRAG_OVERLAP=0): the retrievers run one after the other"),
                      "stage_times": "roofline.avg_launch_ms / bm25.avg_ms_per_step: each retriever alone, in a serial step"},
           "roofline": roofline, "e2e": e2e, "latency": latency, "gpu_launches": launches_per_step * a.steps,
           "clocks": clock_info, "planted_top1_in_top10": hit, "c2": c2,
           "chroma_hnsw_recall_at_10": None,
           "chroma_note": "chromadb/hnswlib are not installable offline: recall of the reference's ANN path vs exact "
                          "search cannot be measured here; this implementation is exact (recall 1.0 by construction)"}
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(a)
    elif rank == 0:
        out["cpu_baseline"] = None
    if world > 1:
        dist.barrier()
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        # The captured CUDA graphs hold NCCL kernels of this communicator; tearing the
        # process group down under them can block forever.  Drop the graphs, drain the
        # device, and leave without destroy_process_group (exit code 0 on every rank).
        del gs, g1, gv
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
