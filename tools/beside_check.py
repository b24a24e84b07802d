"""Do the dense call and the BM25 call overlap when they are enqueued on two streams?  Times, with CUDA
events, at 10M x 768, batch B: the dense call alone, the BM25 call alone, and both on two streams (BM25
launched first / dense launched first).   python tools/beside_check.py [rows] [batch]
Round-2 result (batch 128): dense 2.34 ms, BM25 1.19 ms, both 3.63 / 3.65 ms = the sum: the two persistent
one-CTA-per-SM kernels (192 KB / 193 KB of shared memory, all of TMEM each) take the SMs in turn.  An
experiment that ran the posting scatter as a small persistent grid sized to fit beside the dense CTAs (2 per
SM, 25 KB of shared memory, maximal carve-out) did not overlap either (3.78 - 3.96 ms) and is not kept."""
import json
import sys

import torch

from classmate_rag_b200 import lexical, ops, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
b = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = torch.device("cuda")
emb = synth.dense_corpus(n, 768, dev)
doc_ptr, tokens = synth.lexical_corpus(n, 30000, 64, dev)
lex = lexical.build_lexical_index(doc_ptr, tokens, 30000)
del doc_ptr, tokens
q = ops.f32_to_bf16(synth.dense_queries(n, 768, b, dev)[0])
qt, qp = lexical.pack_queries(synth.lexical_queries(b, 30000))
qt, qp = qt.to(dev), qp.to(dev)
ws = ops.DenseWorkspace(n, 768, b, 24, dev)
import ctypes as C
from classmate_rag_b200 import _lib
st = lex.struct()
buf = ops.TopkBuffers(b, 10, _lib.load().cmr_bm25_workspace_bytes(C.byref(st), b, 10), dev)
main, side = torch.cuda.Stream(), torch.cuda.Stream()


def dense():
    ops.dense_topk(emb, q, 24, workspace=ws)


def bm25(algo):
    ops.bm25_topk(lex, qt, qp, 10, buffers=buf, algo=algo)


def timed(fn, iters=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def both(algo, bm_first):
    cur = torch.cuda.current_stream()
    main.wait_stream(cur)
    side.wait_stream(cur)
    if bm_first:
        with torch.cuda.stream(side):
            bm25(algo)
        with torch.cuda.stream(main):
            dense()
    else:
        with torch.cuda.stream(main):
            dense()
        with torch.cuda.stream(side):
            bm25(algo)
    cur.wait_stream(main)
    cur.wait_stream(side)


out = {"rows": n, "batch": b, "dense_ms": timed(dense), "bm25_auto_ms": timed(lambda: bm25("auto"))}
for algo in ("auto",):
    for first in (True, False):
        out[f"both_{algo}_{'bm25' if first else 'dense'}_first_ms"] = timed(lambda: both(algo, first))
print(json.dumps(out))
