"""N > 1 host logic on CPU: world_size-2 gloo processes shard a list of images, build packed detection
records and all-gather them; the gathered result must equal the single-process result."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

STEPS, NC, MAXDET = 4, 7, 5


def _fake_record(img_idx: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(100 + img_idx)
    k = img_idx % (MAXDET + 1)
    rec = torch.zeros(MAXDET, 10 + STEPS * NC)
    rec[:k, 0] = 1
    rec[:k, 1:] = torch.rand(k, 9 + STEPS * NC, generator=g)
    return rec


def _worker(rank, world, port, n_images, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from glass_text_spotting_b200 import parallel
    b, e = parallel.shard_range(n_images, rank, world)
    assert e - b == n_images // world
    rec = torch.stack([_fake_record(i) for i in range(b, e)])
    gathered = parallel.all_gather_records(rec)
    if rank == 0:
        q.put(gathered.clone())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_everything():
    from glass_text_spotting_b200 import parallel
    for n in (0, 1, 7, 8, 33):
        for world in (1, 2, 3, 8):
            r = [parallel.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in r]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gather_equals_single_process():
    from glass_text_spotting_b200 import parallel
    world, n_images = 2, 6
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_images, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    want = torch.stack([_fake_record(i) for i in range(n_images)]).view(world, n_images // world, MAXDET, -1)
    assert torch.equal(got, want)
    dets = parallel.unpack_detections(got, STEPS, NC)
    assert [len(d["scores"]) for d in dets] == [i % (MAXDET + 1) for i in range(n_images)]
    assert dets[3]["pred_text_prob"].shape == (3, STEPS, NC)
    # rank 0's consumer: the gathered tensor -> Instances -> the evaluator's records
    from glass_text_spotting_b200 import evaluation
    from glass_text_spotting_b200.text import TextDecoder
    insts = evaluation.instances_from_packed(got.reshape(n_images, MAXDET, -1), [(64, 64)] * n_images, STEPS, NC)
    assert [len(i) for i in insts] == [len(d["scores"]) for d in dets]
    assert all(torch.equal(i.pred_boxes.tensor, d["pred_boxes"]) for i, d in zip(insts, dets))
    dec = TextDecoder("abcde")   # 2 + 5 = 7 classes
    recs = [evaluation.instances_to_coco_json(inst, img_id, dec) for img_id, inst in enumerate(insts)]
    assert all(r["image_id"] == i for i, rr in enumerate(recs) for r in rr)
    assert sum(len(r) for r in recs) <= sum(len(i) for i in insts)


def test_single_process_gather_is_identity():
    from glass_text_spotting_b200 import parallel
    rec = torch.stack([_fake_record(i) for i in range(3)])
    assert torch.equal(parallel.all_gather_records(rec)[0], rec)


def _ragged_worker(rank, world, port, n_images, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from glass_text_spotting_b200 import parallel
    b, e = parallel.shard_range(n_images, rank, world)
    rec = torch.stack([_fake_record(i) for i in range(b, e)]) if e > b else torch.zeros(0, MAXDET, 10 + STEPS * NC)
    out = parallel.gather_sharded(rec, n_images)
    q.put((rank, out.clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_uneven_shards_gather_in_dataset_order():
    """7 images on 3 ranks (3 + 2 + 2) and 2 images on 3 ranks (1 + 1 + 0): one all-gather of padded records, every rank
    ends with the whole dataset in order (the reference: comm.gather of pickled lists, text_evaluator.py:246-252)."""
    from glass_text_spotting_b200 import parallel
    for n_images in (7, 2):
        world = 3
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        procs = [ctx.Process(target=_ragged_worker, args=(r, world, port, n_images, q)) for r in range(world)]
        for p in procs:
            p.start()
        got = dict(q.get(timeout=120) for _ in range(world))
        for p in procs:
            p.join(timeout=120)
            assert p.exitcode == 0
        want = torch.stack([_fake_record(i) for i in range(n_images)])
        for r in range(world):
            assert torch.equal(got[r], want), (n_images, r)
    # single process: identity
    rec = torch.stack([_fake_record(i) for i in range(4)])
    assert torch.equal(parallel.gather_sharded(rec, 4), rec)
