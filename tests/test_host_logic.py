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
        # the peer-store transport needs CUDA + NCCL: under gloo the factory reports "unavailable" and callers keep the collective
        ok3 = ok3 and sh.PeerGather.create((3, 3, 4, 4), torch.float32, "cpu") is None
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


# ---- round 2: fake (meta) kernels, autograd registration, fused-caller gating - all without a GPU ----
def test_fake_kernels_give_reference_shapes():
    import torch
    from torch._subclasses.fake_tensor import FakeTensorMode

    vision_b200._lib.load_ops()
    ops = torch.ops.vision_b200
    with FakeTensorMode():
        x = torch.empty(2, 50, 20, 30, device="cuda")
        r = torch.empty(7, 5, device="cuda")
        assert ops.roi_align(x, r, 0.5, 7, 5, 2, False).shape == (7, 50, 7, 5)
        o, a = ops.roi_pool(x, r, 0.5, 3, 3)
        assert o.shape == (7, 50, 3, 3) and a.dtype == torch.int32
        o, m = ops.ps_roi_align(x, r, 0.5, 5, 5, 2)
        assert o.shape == (7, 2, 5, 5) and m.dtype == torch.int32
        o, m = ops.ps_roi_pool(x, r, 0.5, 5, 5)
        assert o.shape == (7, 2, 5, 5)
        assert ops._roi_align_backward(torch.empty(7, 50, 7, 5, device="cuda"), r, 0.5, 7, 5, 2, 50, 20, 30, 2, False).shape == (2, 50, 20, 30)
        assert ops.resize(x, 9, 11, 0, True).shape == (2, 50, 9, 11)
        w = torch.empty(8, 50, 3, 3, device="cuda")
        off = torch.empty(2, 18, 20, 30, device="cuda")
        msk = torch.empty(2, 9, 20, 30, device="cuda")
        out = ops.deform_conv2d(x, w, off, msk, torch.empty(8, device="cuda"), 1, 1, 1, 1, 1, 1, 1, 1, True)
        assert out.shape == (2, 8, 20, 30)
        feats = [torch.empty(2, 16, 40 // s, 48 // s, device="cuda") for s in (1, 2, 4)]
        o, lv = ops.multiscale_roi_align(feats, r, [0.25, 0.125, 0.0625], 7, 7, 2, 2, 4, 224.0, 4.0, 1e-6)
        assert o.shape == (7, 16, 7, 7) and lv.shape == (7,) and lv.dtype == torch.int32


def test_fused_caller_gating_and_crop_arithmetic():
    import torch
    from vision_b200 import ops as vops, transforms as vtf
    from torchvision.transforms import InterpolationMode

    # CPU tensors never take the fused paths (install() then calls the reference body)
    feats = [torch.rand(1, 4, 32 // s, 32 // s) for s in (1, 2)]
    assert not vops.multiscale_roi_align_supported(feats, [torch.rand(3, 4)], (7, 7), 2)
    assert not vtf.classification_preprocess_supported(torch.zeros(3, 300, 400, dtype=torch.uint8), [224], [256], InterpolationMode.BILINEAR, True)
    # output-size rule of the preset: shorter edge to resize_size, then the centred crop (transforms/functional.py:353-384, center_crop)
    assert vtf.compute_resized_output_size((375, 500), size=[256]) == [256, 341]
    assert vtf._crop_hw([224]) == (224, 224) and vtf._crop_hw((200, 210)) == (200, 210)
    with pytest.raises(RuntimeError, match="unsupported input"):
        vtf.classification_preprocess(torch.zeros(3, 300, 400), [224], [256], (0.5,) * 3, (0.5,) * 3)
    from vision_b200 import detection
    assert not detection._fusable(torch.rand(4, 4))


def test_install_rebinds_and_restores_python_entry_points():
    import torchvision
    from torchvision.models.detection import roi_heads, rpn
    from torchvision.ops import poolers
    from torchvision.transforms import _presets

    before = (torchvision.ops.boxes.batched_nms, poolers._multiscale_roi_align, roi_heads.RoIHeads.postprocess_detections,
              rpn.RegionProposalNetwork.filter_proposals, _presets.ImageClassification.forward)
    vision_b200.install()
    try:
        during = (torchvision.ops.boxes.batched_nms, poolers._multiscale_roi_align, roi_heads.RoIHeads.postprocess_detections,
                  rpn.RegionProposalNetwork.filter_proposals, _presets.ImageClassification.forward)
        assert all(a is not b for a, b in zip(before, during))
        # CPU inputs still run the reference bodies through the rebinding
        import torch
        b = torch.tensor([[0.0, 0, 10, 10], [1, 1, 11, 11], [50, 50, 60, 60]])
        keep = torchvision.ops.batched_nms(b, torch.tensor([0.9, 0.8, 0.7]), torch.tensor([0, 0, 1]), 0.5)
        assert keep.tolist() == [0, 2]
    finally:
        vision_b200.uninstall()
    after = (torchvision.ops.boxes.batched_nms, poolers._multiscale_roi_align, roi_heads.RoIHeads.postprocess_detections,
             rpn.RegionProposalNetwork.filter_proposals, _presets.ImageClassification.forward)
    assert all(a is b for a, b in zip(before, after))


def test_peer_gather_slot_arithmetic():
    """PeerGather hands the kernels this rank's slot in every rank's buffer: the local copy first, then the peers in rank order, in
    the buffer the step uses (two alternating buffers); checked with a stand-in handle (no GPU, no process group)."""
    import torch

    class Handle:
        buffer_ptrs = [0x1000_0000, 0x2000_0000, 0x3000_0000]
        multicast_ptr = 0x9000_0000

        def barrier(self, channel=0):
            self.calls = getattr(self, "calls", 0) + 1

    world, rank, shard = 3, 1, (4, 5)
    buf = torch.zeros((2, world) + shard, dtype=torch.float16)
    pg = sharded.PeerGather(buf, Handle(), rank, world)
    shard_bytes, half = 4 * 5 * 2, 3 * 4 * 5 * 2
    assert pg.shard_shape == shard
    for step in range(4):
        pg.advance()
        k = step % 2                      # the first step uses buffer 0
        assert pg.cur == k
        own = rank * shard_bytes + k * half
        assert pg.dst_ptrs == [0x2000_0000 + own, 0x1000_0000 + own, 0x3000_0000 + own]
        assert pg.mc_ptr == 0x9000_0000 + own
        assert pg.gathered().shape == (world * 4, 5) and pg.gathered().data_ptr() == buf[k].data_ptr()
    pg.barrier()
    assert pg.hdl.calls == 1
