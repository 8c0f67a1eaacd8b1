"""CPU, world_size 2 over gloo: the multi-GPU partition of the path (lerf_pytorch_b200/sharding.py).

Each rank computes ITS share with the oracle standing in for the kernels (tests may use the oracle; the product
never does): per-image shards, and row bands of one frame computed from an input crop of only the rows
``band_input_rows`` says the band needs.  Rank 0 gathers the pieces and compares them with the single-process
result: equal means the partition (incl. the 7-row halo) is exact and nothing has to be exchanged between ranks.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import lerf_oracle as orc
        import util
        from lerf_pytorch_b200 import sharding
        orc.set_threads(2)
        luts = orc.load_luts(util.lut_dir("lerf-g"), linear=False)
        S = 4

        # ---- per-image sharding: 5 images over 2 ranks, results gathered by rank 0
        imgs = [util.uniform_image(700 + i, 24, 31) for i in range(5)]
        mine = sharding.image_shard(len(imgs), rank, world)
        outs = {i: orc.to_uint8_hwc(orc.lerf_sr(imgs[i], luts, S, S)[0]) for i in mine}
        gathered = [None] * world
        dist.all_gather_object(gathered, outs)
        ok_img = True
        if rank == 0:
            merged = {}
            for d in gathered:
                merged.update(d)
            ok_img = sorted(merged) == list(range(len(imgs))) and all(
                np.array_equal(merged[i], orc.to_uint8_hwc(orc.lerf_sr(imgs[i], luts, S, S)[0])) for i in range(len(imgs)))

        # ---- row bands of one frame: each rank only looks at the input rows its band needs
        img = util.natural_image(91, 75, 40)
        H = img.shape[0]
        oH = S * H
        bands = sharding.row_bands(oH, world, align=S)
        oy0, oy1 = bands[rank]
        r0, r1, _, _ = sharding.band_input_rows(H, oH, S, oy0, oy1)
        crop = np.ascontiguousarray(img[r0:r1])
        band_full, _, _ = orc.lerf_sr(crop, luts, S, S)  # float64 [3, S*(r1-r0), oW]; output row o of the frame = row o - S*r0
        piece = band_full[:, oy0 - S * r0:oy1 - S * r0]
        pieces = [None] * world
        dist.all_gather_object(pieces, (oy0, oy1, piece))
        ok_band, ok_cover = True, True
        if rank == 0:
            ref, _, _ = orc.lerf_sr(img, luts, S, S)
            rows = []
            for a, b, pc in sorted(pieces, key=lambda t: t[0]):
                rows.append((a, b))
                ok_band = ok_band and np.array_equal(pc, ref[:, a:b])
            ok_cover = rows[0][0] == 0 and rows[-1][1] == oH and all(rows[i][1] == rows[i + 1][0] for i in range(len(rows) - 1))

        # ---- timing reduction the bench uses: max over ranks
        t = torch.tensor([10.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            q.put({"img": bool(ok_img), "band": bool(ok_band), "cover": bool(ok_cover), "tmax": float(t.item()),
                   "halo": (r0, r1)})
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_partition_is_exact():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res["img"], "per-image shards differ from the single-process result"
    assert res["cover"], "row bands do not tile the output"
    assert res["band"], "row bands computed from cropped inputs differ from the full-frame result"
    assert res["tmax"] == 11.0


def test_partition_helpers():
    from lerf_pytorch_b200 import sharding
    assert sharding.image_shard(5, 1, 2) == [1, 3]
    assert sharding.image_shard(0, 0, 8) == []
    with pytest.raises(ValueError):
        sharding.image_shard(3, 2, 2)
    for oH, world, align in ((17280, 8, 8), (300, 8, 4), (5, 8, 1), (96, 3, 4)):
        b = sharding.row_bands(oH, world, align)
        assert len(b) == world and b[0][0] == 0 and b[-1][1] == oH
        assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        assert all(x[0] % align == 0 for x in b)
    # cfg-5: 2160 input rows x8; an interior band needs exactly its rows + 7 on each side
    r0, r1, c0, c1 = sharding.band_input_rows(2160, 17280, 8, 8 * 540, 8 * 1080)
    assert (c0, c1) == (539, 1081) and (r0, r1) == (533, 1087)
    assert sharding.band_halo_rows() == 7
    # image edges clamp
    assert sharding.band_input_rows(100, 400, 4, 0, 40)[0] == 0
    assert sharding.band_input_rows(100, 400, 4, 360, 400)[1] == 100
