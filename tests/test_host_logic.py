"""CPU suite, part 3: host-side mirror logic (output-size rules, interpolation checks, sharding)."""
import os
import socket
import sys

import pytest
import torch

import vision_b200
from vision_b200 import sharded, transforms as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("canvas", [(17, 11), (11, 17), (2160, 3840), (5, 5)])
@pytest.mark.parametrize("size,max_size", [(17, None), ([17], None), ((17,), None), ([12, 13], None), ((12, 13), None),
                                           ([9], 14), (None, 10), ([20], 40)])
def test_output_size_matches_reference(canvas, size, max_size):
    tv = pytest.importorskip("torchvision")
    from torchvision.transforms.v2.functional._geometry import _compute_resized_output_size as ref

    try:
        want = ref(canvas, size=size, max_size=max_size)
    except ValueError as e:
        with pytest.raises(ValueError):
            T.compute_resized_output_size(canvas, size, max_size)
        return
    assert T.compute_resized_output_size(canvas, size, max_size) == want


def test_output_size_errors():
    with pytest.raises(ValueError, match="max_size should only be passed"):
        T.compute_resized_output_size((10, 10), [5, 5], max_size=7)
    with pytest.raises(ValueError, match="strictly greater"):
        T.compute_resized_output_size((10, 20), [8], max_size=8)
    with pytest.raises(ValueError, match="max_size must be an integer"):
        T.compute_resized_output_size((10, 20), None, max_size=None)


def test_interpolation_check():
    assert T._mode_value("bilinear") == "bilinear"
    assert T._mode_value(T.InterpolationMode.BICUBIC) == "bicubic"
    assert T._mode_value(2) == "bilinear" and T._mode_value(3) == "bicubic"
    with pytest.raises(ValueError, match="Invalid interpolation mode"):
        T._mode_value("cubic")
    with pytest.raises(ValueError):
        T._mode_value(2.5)
    tv = pytest.importorskip("torchvision")
    from torchvision.transforms import InterpolationMode as TVMode

    assert T._mode_value(TVMode.BILINEAR) == "bilinear"
    assert not T.supports(torch.zeros(3, 4, 4), TVMode.BILINEAR)          # CPU tensor -> reference path


def test_roi_format_helpers():
    boxes = [torch.rand(3, 4), torch.rand(2, 4)]
    rois = vision_b200.ops.convert_boxes_to_roi_format(boxes)
    assert rois.shape == (5, 5) and rois[:, 0].tolist() == [0, 0, 0, 1, 1]
    with pytest.raises(AssertionError, match="Tensor\\[K, 5\\]"):
        vision_b200.ops.check_roi_boxes_shape(torch.rand(3, 4))
    with pytest.raises(AssertionError, match="List\\[Tensor\\[L, 4\\]\\]"):
        vision_b200.ops.check_roi_boxes_shape([torch.rand(3, 5)])


def test_shard_bounds():
    for n in (0, 1, 7, 8, 1000, 1024):
        for world in (1, 2, 3, 8):
            spans = [sharded.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    from vision_b200 import sharded as sh

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        # equal-size outputs: a stand-in op (x * 2) over this rank's image shard
        imgs = torch.arange(6 * 3 * 4 * 4, dtype=torch.float32).reshape(6, 3, 4, 4)
        s, e = sh.shard_bounds(6, rank, world)
        out = sh.sharded_apply(lambda t: t * 2, [imgs[s:e]])
        ok1 = torch.equal(out, imgs * 2)
        # variable-length keep lists: stand-in "nms" keeps the even indices
        probs = [(torch.zeros(5 + rank + j, 4), torch.zeros(5 + rank + j), torch.zeros(5 + rank + j, dtype=torch.int64))
                 for j in range(2)]
        fake = lambda b, sc, i, thr: torch.arange(0, b.shape[0], 2, dtype=torch.int64)
        res = sh.sharded_batched_nms(fake, probs, 0.5, capacity=64)
        ok2 = len(res) == world and all(
            torch.equal(res[r][j], torch.arange(0, 5 + r + j, 2, dtype=torch.int64)) for r in range(world) for j in range(2))
        # chunked gather (the overlapped path; no side stream on CPU): rank-major result, lazily packed
        og = sh.OverlappedGather()
        mine = imgs[s:e]
        lazy = sh.sharded_apply_overlapped(lambda t: t * 2, mine, chunks=3, gather=og)
        ok3 = tuple(lazy.shape) == tuple(imgs.shape) and torch.equal(lazy.materialize(), imgs * 2) and \
            torch.equal(lazy.rank(1 - rank).flatten(0, 1), (imgs * 2)[(1 - rank) * 3:(2 - rank) * 3])
        one = sh.sharded_apply_overlapped(lambda t: t + 1, mine, chunks=1, gather=og)
        ok3 = ok3 and torch.equal(one, imgs + 1)
        q.put((rank, ok1, ok2 and ok3))
    finally:
        dist.destroy_process_group()


def test_sharded_gather_gloo_world2():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=90) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in results) == [0, 1]
    assert all(r[1] and r[2] for r in results)
