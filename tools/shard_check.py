"""Multi-GPU check of the sharded hot path (run under torchrun, one rank per GPU):
peer-memory exchange (rows inside the messages, and rows pulled from their owners' shared
matrices) == NCCL all-gather exchange == unsharded search, bit for bit, eager and from a CUDA graph.  Prints "shard_check ok" on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from classmate_rag_b200 import lexical, ops, sharding, synth  # noqa: E402
from classmate_rag_b200.engine import GraphedSearch, HybridEngine, PipelinedSearch, SearchParams  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n, d, vocab, nq = 60_000, 256, 3000, 12
    p = SearchParams(top_k=10)
    # unsharded truth (every rank computes it: small)
    emb = synth.dense_corpus(n, d, dev)
    doc_ptr, tokens = synth.lexical_corpus(n, vocab, 24, dev)
    stats = lexical.corpus_stats(doc_ptr, tokens, vocab)
    full = HybridEngine(emb, lexical.build_lexical_index(doc_ptr, tokens, vocab, stats=stats))
    q, _ = synth.dense_queries(n, d, nq, dev)
    terms = synth.lexical_queries(nq, vocab)
    qb = ops.f32_to_bf16(q)
    qt, qp = lexical.pack_queries(terms)
    qt, qp = qt.to(dev), qp.to(dev)
    want = [t.clone() for t in full.search(qb, qt, qp, p)]
    # this rank's shard
    lo, hi = sharding.shard_range(n, rank, world)
    t_lo, t_hi = int(doc_ptr[lo]), int(doc_ptr[hi])
    gstats = sharding.global_corpus_stats(doc_ptr[lo:hi + 1] - doc_ptr[lo], tokens[t_lo:t_hi], vocab, doc_lo=lo,
                                          n_docs_total=n, token_offset=t_lo)
    sh_lex = lexical.build_lexical_index(doc_ptr[lo:hi + 1] - doc_ptr[lo], tokens[t_lo:t_hi], vocab, stats=gstats)
    results = {}
    shared = sharding.shared_rows(hi - lo, d, dev)      # this shard's rows, readable by every rank
    shared.copy_(emb[lo:hi])
    for name, peer in (("nccl", False), ("peer", True), ("peer_pull", True)):
        comm = sharding.ShardComm(peer_memory=peer)
        eng = HybridEngine(shared if name == "peer_pull" else emb[lo:hi].contiguous(), sh_lex, row_offset=lo, comm=comm)
        for it in range(3):      # several steps: epochs / parity buffers of the peer exchange
            got = [t.clone() for t in eng.search(qb, qt, qp, p)]
        torch.cuda.synchronize()
        if peer:
            assert comm.peer is not None, f"peer memory unavailable: {comm.peer_error}"
            assert int(comm.peer.timeout.item()) == 0
            assert comm.peer.pull_rows == (name == "peer_pull"), name
        for a, b in zip(got, want):
            assert a.cpu().numpy().tobytes() == b.cpu().numpy().tobytes(), name
        gs = GraphedSearch(eng, p, nq, max_terms=16)
        assert gs.graph is not None
        qh = q.cpu().numpy()
        for it in range(4):
            out = gs(qh, terms)
        for a, b in zip(out, want):
            assert a.tobytes() == b.cpu().numpy().tobytes(), name + " graph"
        # two steps in flight (two graph objects, two streams, two exchange buffer sets)
        ps = PipelinedSearch(eng, p, nq, max_terms=16)
        assert ps.independent == peer, name
        outs = []
        for it in range(5):
            r = ps.submit(qh, terms)
            if r is not None:
                outs.append(r)
        outs.append(ps.drain())
        assert len(outs) == 5
        for o in outs:
            for a, b in zip(o, want):
                assert a.tobytes() == b.cpu().numpy().tobytes(), name + " pipelined"
        for it in range(4):
            dev_out = ps.launch_resident(q, qt, qp)
        ps.wait()
        torch.cuda.synchronize()
        for a, b in zip(dev_out, want):
            assert a.cpu().numpy().tobytes() == b.cpu().numpy().tobytes(), name + " pipelined resident"
        del ps
        # the other message shapes: no lexical list (non-hybrid), no rows (MMR off)
        for hyb, mmr in ((False, True), (True, False), (False, False)):
            p2 = SearchParams(top_k=10, hybrid=hyb, use_mmr=mmr)
            w2 = [t.clone() for t in full.search(qb, qt, qp, p2)]
            g2 = [t.clone() for t in eng.search(qb, qt, qp, p2)]
            torch.cuda.synchronize()
            for a, b in zip(g2, w2):
                assert a.cpu().numpy().tobytes() == b.cpu().numpy().tobytes(), (name, hyb, mmr)
        results[name] = True
        del gs, eng
        torch.cuda.synchronize()
        dist.barrier()
    if rank == 0:
        print("shard_check ok", world, "ranks", results, flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
