"""N>1 path on CPU: two gloo ranks shard the nuclei with nfx_partition (contiguous ranges aligned to
batch_size, no data-path collective), each rank runs the per-range pipeline (the oracle stands in
for the kernels here -- tests may use it), results are merged in input order and must equal the
single-process run bit for bit, including the batch-coupled mean_h column."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    for p in (os.path.join(ROOT, "nuclei-feature-extraction_b200"), os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    import nfx
    import nfx_oracle as o
    from nfx import synth
    torch.set_num_threads(1)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    tile = synth.synth_tile(256, 256, 9)
    xy, off = synth.synth_polygons(70, 256, 256, 9)
    rings = synth.rings_of(xy, off)
    B = 16
    bounds = nfx.partition(len(rings), B, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    keys, cents, feats, names = o.extract(rings[lo:hi], tile, ["color"], 64, B)
    # merge on rank 0 in input order (the product does this with D2H copies at the range offset)
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, keys, feats))
    if rank == 0:
        full = np.zeros((len(rings), feats.shape[1]))
        allkeys = [None] * len(rings)
        for lo_, hi_, k_, f_ in gathered:
            full[lo_:hi_] = f_
            allkeys[lo_:hi_] = k_
        wk, wc, want, _ = o.extract(rings, tile, ["color"], 64, B)
        q.put((allkeys == wk, bool(np.array_equal(full, want, equal_nan=True)), bounds))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    keys_ok, feats_ok, bounds = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert keys_ok and feats_ok
    assert bounds[1] % 16 == 0
