"""N > 1 path on CPU: world_size-2 gloo run of the result-record gather that bench.py performs once per
step over NCCL (batch-sharded images, no data-path collective)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


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
    from rsprompter_b200.results import gather_mask_logits, gather_records, pack_records, unpack_records
    B, M = 3, 5
    g = torch.Generator().manual_seed(rank)
    boxes = torch.rand(B, M, 4, generator=g)
    scores = torch.rand(B, M, generator=g)
    labels = torch.randint(0, 10, (B, M), generator=g)
    counts = torch.tensor([M, 2, 0], dtype=torch.int32) + rank
    counts = counts.clamp(max=M)
    rec = pack_records(boxes, scores, labels)
    all_rec, all_cnt = gather_records(rec, counts)
    assert all_rec.shape == (world * B, M, 6) and all_cnt.shape == (world * B,)
    mine = unpack_records(all_rec[rank * B:(rank + 1) * B], all_cnt[rank * B:(rank + 1) * B])
    ok = all(torch.equal(mine[i]["bboxes"], boxes[i, :counts[i]]) and
             torch.equal(mine[i]["labels"], labels[i, :counts[i]]) for i in range(B))
    logits = torch.randn(B * M, 8, 8, generator=g)
    all_logits = gather_mask_logits(logits)
    ok = ok and all_logits.dtype == torch.float16 and all_logits.shape == (world * B * M, 8, 8)
    ok = ok and torch.equal(all_logits[rank * B * M:(rank + 1) * B * M], logits.half())
    # the full record (bit-packed masks + rows + counts) in ONE collective
    import numpy as np
    from rsprompter_b200.results import ResultRecord
    H, W = 16, 24
    rr = ResultRecord(B, M, (H, W), device="cpu")
    masks = torch.rand(B, M, H, W, generator=g) > 0.5
    rr.mask_bits.copy_(torch.from_numpy(np.packbits(masks.numpy(), axis=-1, bitorder="little")))
    rr.rows.copy_(rec)
    rr.counts.copy_(counts)
    parts = rr.split(rr.gather())
    ok = ok and len(parts) == world and torch.equal(parts[rank].buf, rr.buf)
    inst = parts[rank].instances()
    ok = ok and all(torch.equal(inst[i]["masks"], masks[i, :counts[i]]) and
                    torch.equal(inst[i]["bboxes"], boxes[i, :counts[i]]) for i in range(B))
    other = parts[1 - rank].counts.tolist()
    ok = ok and other == (torch.tensor([M, 2, 0]) + (1 - rank)).clamp(max=M).tolist()
    q.put((rank, ok, all_cnt.tolist()))
    dist.destroy_process_group()


def test_record_gather_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert res[0][2] == res[1][2]          # every rank sees the same global counts


def test_coco_rle_of_record_masks():
    """mask_to_coco_rle restates pycocotools' published format (maskApi.c rleEncode + rleToString); hand-derived
    known answers, the inverse, and the record -> CocoMetric 'segm' result dicts path."""
    import numpy as np
    import torch
    from rsprompter_b200.results import ResultRecord, coco_rle_to_mask, mask_to_coco_rle, record_to_coco_results
    m = np.array([[0, 1], [1, 1]], dtype=bool)                       # column-major 0,1,1,1 -> runs [1, 3]
    assert mask_to_coco_rle(m) == dict(size=[2, 2], counts=b"13")
    assert mask_to_coco_rle(np.ones((4, 5), dtype=bool))["counts"] == b"0d0"     # [0, 20]: 20 = 0x14 -> 'd', '0'
    assert mask_to_coco_rle(np.zeros((4, 5), dtype=bool))["counts"] == b"d0"
    g = np.random.default_rng(0)
    for shape in ((1, 1), (7, 13), (64, 40), (300, 257)):
        for p in (0.02, 0.5, 0.97):
            mk = g.random(shape) < p
            mk[: shape[0] // 3] = mk[:1]                                  # long runs -> negative count differences
            assert np.array_equal(coco_rle_to_mask(mask_to_coco_rle(mk)), mk)
    rec = ResultRecord(2, 3, (16, 24))
    masks = torch.from_numpy(g.random((2, 3, 16, 24)) < 0.4)
    bits = np.packbits(masks.numpy(), axis=-1, bitorder="little")
    rec.mask_bits.copy_(torch.from_numpy(bits))
    rec.rows.copy_(torch.tensor([[[1., 2., 11., 22., 0.9, 4.]] * 3] * 2))
    rec.counts.copy_(torch.tensor([2, 0], dtype=torch.int32))
    res = record_to_coco_results(rec, image_ids=[17, 18], label_to_cat={4: 5})
    assert len(res) == 2 and res[0]["image_id"] == 17 and res[0]["category_id"] == 5 and res[0]["bbox"] == [1., 2., 10., 20.]
    assert np.array_equal(coco_rle_to_mask(res[1]["segmentation"]), masks[0, 1].numpy())
