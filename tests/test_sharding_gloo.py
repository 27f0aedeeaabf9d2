"""Host-side logic of the multi-GPU path on CPU: nnz-balanced document shards and the reduction
protocol (sum of per-rank statistics / ELBO scalars == whole-corpus result), run as a real
world_size-2 job over the gloo backend.  The per-rank "device" is played by the oracle here (this is
a test of the sharding / reduction host logic, not of the CUDA kernels)."""
import os
import tempfile

import numpy


def test_shard_bounds_cover_and_balance():
    from pylda_b200 import native, synthetic
    row_ptr, ids, cts = synthetic.synthetic_corpus(500, 800, seed=4, length="zipf")
    for n_ranks in (1, 2, 3, 8):
        b = native.shard_bounds(row_ptr, n_ranks)
        assert b[0] == 0 and b[-1] == 500 and len(b) == n_ranks + 1
        assert numpy.all(numpy.diff(b) >= 0)
        nnz = numpy.diff(row_ptr[b])
        assert nnz.sum() == len(ids)
        longest = int(numpy.diff(row_ptr).max())
        assert nnz.max() - nnz.min() <= 2 * longest + 1          # balanced to within a document or two
        pieces = [native.shard_csr(row_ptr, ids, cts, int(b[r]), int(b[r + 1])) for r in range(n_ranks)]
        assert numpy.array_equal(numpy.concatenate([p[1] for p in pieces]), ids)
        assert numpy.array_equal(numpy.concatenate([p[2] for p in pieces]), cts)
        for p in pieces:
            assert p[0][0] == 0 and p[0][-1] == len(p[1])


def test_shard_bounds_edge_cases():
    from pylda_b200 import native
    assert list(native.shard_bounds(numpy.array([0]), 4)) == [0, 0, 0, 0, 0]            # empty corpus
    assert list(native.shard_bounds(numpy.array([0, 5]), 2))[-1] == 1                    # one document, two ranks
    b = native.shard_bounds(numpy.array([0, 1, 2, 3]), 8)                                # more ranks than documents
    assert b[0] == 0 and b[-1] == 3 and numpy.all(numpy.diff(b) >= 0)


def _rank_main(rank, world, init_file, out_dir):
    import torch
    import torch.distributed as dist
    from oracle import estep_oracle as O
    from pylda_b200 import native, synthetic
    dist.init_process_group("gloo", init_method="file://" + init_file, rank=rank, world_size=world)
    K, V = 6, 300
    row_ptr, ids, cts = synthetic.synthetic_corpus(60, V, seed=21, length="poisson", mean_len=40)
    eta = synthetic.initial_eta(K, V, 0)
    alpha = numpy.full(K, 1.0 / K)
    b = native.shard_bounds(row_ptr, world)
    rp, ii, cc = native.shard_csr(row_ptr, ids, cts, int(b[rank]), int(b[rank + 1]))
    r = O.e_step(rp, ii, cc, eta, alpha, 50, 1e-6)
    # the packed buffer the library all-reduces: [phi_ss | doc_ll | D_local]
    packed = torch.from_numpy(numpy.concatenate([r["phi_ss"].reshape(-1), [r["doc_ll"], float(len(rp) - 1)]]))
    dist.all_reduce(packed, op=dist.ReduceOp.SUM)
    gathered = [None] * world
    dist.all_gather_object(gathered, r["gamma"])
    if rank == 0:
        numpy.savez(os.path.join(out_dir, "out.npz"), packed=packed.numpy(), gamma=numpy.concatenate(gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_equals_single_process():
    import torch.multiprocessing as mp
    from oracle import estep_oracle as O
    from pylda_b200 import synthetic
    world = 2
    with tempfile.TemporaryDirectory() as tmp:
        init_file = os.path.join(tmp, "rendezvous")
        mp.spawn(_rank_main, args=(world, init_file, tmp), nprocs=world, join=True)
        z = numpy.load(os.path.join(tmp, "out.npz"))
    K, V = 6, 300
    row_ptr, ids, cts = synthetic.synthetic_corpus(60, V, seed=21, length="poisson", mean_len=40)
    ref = O.e_step(row_ptr, ids, cts, synthetic.initial_eta(K, V, 0), numpy.full(K, 1.0 / K), 50, 1e-6)
    phi = z["packed"][:-2].reshape(K, V)
    assert numpy.allclose(phi, ref["phi_ss"], rtol=1e-12, atol=1e-300)
    assert abs(z["packed"][-2] - ref["doc_ll"]) <= 1e-12 * abs(ref["doc_ll"])
    assert z["packed"][-1] == 60.0
    assert numpy.array_equal(z["gamma"], ref["gamma"])      # per-document results do not depend on the shard
