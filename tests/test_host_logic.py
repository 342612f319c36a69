"""CPU tests of the host-side logic: weight packing into tap form, ConvTranspose phase decomposition, merged heads,
state_dict interchange with the oracle, C-ABI export list, argument validation (no kernels are launched)."""
import ctypes
import os
import re

import pytest
import torch
import torch.nn.functional as F

from opticalflowscivis_b200 import _C, ifnet
from oracle.ifnet_ref import IFNetRef
from tap_eval import run_layer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cl(x):      # NC(D)HW -> [N][D][H][W][C] (D = 1 for 2-D)
    if x.dim() == 4:
        x = x.unsqueeze(2)
    return x.permute(0, 2, 3, 4, 1).contiguous()


def _pad_c(x, c):
    return F.pad(x, (0, c - x.shape[-1]))


@pytest.mark.parametrize("nd", [2, 3])
def test_conv_tap_form_matches_torch(nd):
    torch.manual_seed(0)
    for (cin, cout, k, s, p) in ((5, 16, 3, 1, 1), (9 if nd == 2 else 11, 32, 3 if nd == 2 else 4, 2, 1)):
        m = ifnet._ConvParams(nd, cin, cout, k, s, p)
        pr = ifnet._PReLUParams(cout)
        pr.weight.data.uniform_(0.1, 0.4)
        lay = ifnet._pack_conv(m, pr)
        x = torch.randn((2, cin) + ((8,) * nd))
        ref = (F.conv2d if nd == 2 else F.conv3d)(x, m.weight, m.bias, stride=s, padding=p)
        ref = F.prelu(ref, pr.weight)
        got = run_layer(lay, _pad_c(_cl(x), lay.cin_s))
        assert torch.allclose(got[..., :cout], _cl(ref), atol=1e-4), (nd, cin, cout)


@pytest.mark.parametrize("nd", [2, 3])
def test_strided_conv_space_to_depth_form_matches_torch(nd):
    """Conv(k=3|4, s=2, p=1) == stride-1 {0,1}^nd-tap conv over the shifted space-to-depth input (ifnet._pack_conv_s2d)."""
    torch.manual_seed(5)
    cin, cout, k = (9 if nd == 2 else 11), 32, (3 if nd == 2 else 4)
    m = ifnet._ConvParams(nd, cin, cout, k, 2, 1)
    pr = ifnet._PReLUParams(cout)
    pr.weight.data.uniform_(0.1, 0.4)
    lay = ifnet._pack_conv_s2d(m, pr, 16, out_s2d=False)
    assert lay.ntaps == 2 ** nd and lay.cin_s == 16 * 2 ** nd and lay.in_stride == 1
    x = torch.randn((2, cin) + ((12,) * nd))
    ref = F.prelu((F.conv2d if nd == 2 else F.conv3d)(x, m.weight, m.bias, stride=2, padding=1), pr.weight)
    xs = ifnet.s2d_shift_pack(_pad_c(_cl(x), 16), nd)
    assert list(xs.shape[1:]) == ([1] if nd == 2 else [7]) + [7, 7, 16 * 2 ** nd]
    got = run_layer(lay, xs)
    assert torch.allclose(got[..., :cout], _cl(ref), atol=1e-4)
    # chained: conv0.0 (s2d out) -> conv0.1 (s2d in) as in IFBlock
    m1 = ifnet._ConvParams(nd, cout, 2 * cout, k, 2, 1)
    lay1 = ifnet._pack_conv_s2d(m1, None, lay.cout_s, out_s2d=False)
    ref1 = (F.conv2d if nd == 2 else F.conv3d)(ref, m1.weight, m1.bias, stride=2, padding=1)
    got1 = run_layer(lay1, ifnet.s2d_shift_pack(got, nd))
    assert torch.allclose(got1[..., :2 * cout], _cl(ref1), atol=2e-4)


@pytest.mark.parametrize("nd", [2, 3])
def test_conv_transpose_phase_form_matches_torch(nd):
    torch.manual_seed(1)
    cin, cout = 6, 5
    m = ifnet._ConvParams(nd, cin, cout, 4, 2, 1, transposed=True)
    lay = ifnet._pack_convT(nd, m.weight.detach(), m.bias.detach(), None, 8, True)
    x = torch.randn((2, cin) + ((5,) * nd))
    ref = (F.conv_transpose2d if nd == 2 else F.conv_transpose3d)(x, m.weight, m.bias, stride=2, padding=1)
    got = run_layer(lay, _pad_c(_cl(x), lay.cin_s))
    assert got.shape[1:4] == _cl(ref).shape[1:4]
    assert torch.allclose(got[..., :cout], _cl(ref), atol=1e-4)


@pytest.mark.parametrize("nd", [2, 3])
def test_depth_to_space_heads_equal_phase_form(nd):
    """The one-conv depth-to-space packing of the final ConvTranspose heads equals the 2^nd-phase packing."""
    torch.manual_seed(3)
    blk = ifnet.IFBlock(nd, 5 + 2 * nd, 32)
    L = blk.layers()
    x = torch.randn((2,) + ((1, 6, 7) if nd == 2 else (4, 6, 7)) + (32,))
    a = run_layer(L[11], x)
    b = run_layer(blk._heads_shuffle, x)
    assert a.shape == b.shape and torch.allclose(a, b, atol=1e-5)


@pytest.mark.parametrize("nd", [2, 3])
def test_block_layers_match_oracle_block(nd):
    """The 12 packed layers (merged conv1.0‖conv2.0, block-diagonal heads, residual pairs) evaluated in tap form on CPU
    reproduce the oracle IFBlock's flow/mask head outputs."""
    torch.manual_seed(2)
    c, cin = 32, 5 + 2 * nd
    blk = ifnet.IFBlock(nd, cin, c)
    from oracle.ifnet_ref import IFBlockRef
    ref = IFBlockRef(nd, cin, c)
    ref.load_state_dict(blk.state_dict())
    x = torch.randn((1, cin) + ((16,) * nd))
    with torch.no_grad():
        h = ref.conv0(x)
        for i in range(4):
            h = getattr(ref, f"convblock{i}")(h) + h
        rflow, rmask = ref.conv1(h), ref.conv2(h)
        L = blk.layers()
        y, skip = _pad_c(_cl(x), 16), None
        for li, lay in enumerate(L):
            inp = y
            y = run_layer(lay, inp, skip if lay.residual else None)
            if 2 <= li <= 9 and li % 2 == 0:
                skip = inp
    nf = 2 * nd
    assert torch.allclose(y[..., :nf], _cl(rflow), atol=2e-4)
    assert torch.allclose(y[..., nf:nf + 1], _cl(rmask), atol=2e-4)


@pytest.mark.parametrize("nd", [2, 3])
def test_state_dict_interchange_and_seeded_init(nd):
    torch.manual_seed(1234)
    mine = ifnet.IFNet(nd)
    torch.manual_seed(1234)
    ref = IFNetRef(nd)
    a, b = mine.state_dict(), ref.state_dict()
    assert list(a) == list(b)
    assert all(torch.equal(a[k], b[k]) for k in a)          # same default init as the reference under the same seed
    ref.load_state_dict(a)                                   # and the keys load both ways
    mine.load_state_dict(b)
    assert sum(p.numel() for p in mine.parameters()) == {2: 3157764, 3: 9101916}[nd]


def test_cabi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "ofsv.h")).read()
    declared = set(re.findall(r"\b(ofsv_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"ofsv_conv_desc"}
    L = _C.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/ofsv.h but not exported by libofsv.so"
    assert declared == set(_C.EXPORTS), declared ^ set(_C.EXPORTS)
    assert L.ofsv_version().startswith(b"ofsv")
    assert ctypes.sizeof(_C.ConvDesc) == 4 * 18 + 4 * _C.MAX_TAPS + 4 * 7


def test_validation_errors_launch_nothing():
    L = _C.lib()
    null = ctypes.c_void_p(0)
    n0 = L.ofsv_launch_count()
    assert L.ofsv_warp3d_f32(null, null, null, null, null, null, 1, 1, 4, 4, 4, 0, null) == _C.EINVAL
    assert b"null" in L.ofsv_last_error()
    assert L.ofsv_corr81_fwd_f32(null, null, null, 1, 0, 4, 4, 0.1, 0, 0, null, null) == _C.EINVAL
    assert L.ofsv_pack_block_input(null, null, null, null, null, null, null, 0, 3, 1, 6, 8, 8, 4, 16, 0, null) == _C.EINVAL
    assert L.ofsv_block_stage_3d(*([null] * 11), 1, 16, 16, 16, 3, 0, 0, 0, 0, null) == _C.EINVAL
    assert L.ofsv_u8_to_f32(null, null, 16, 255.0, null) == _C.EINVAL
    assert L.ofsv_sq_err_f64(null, null, null, null, 1, 16, 1.0, null) == _C.EINVAL
    assert L.ofsv_ssim2d_f64(null, null, null, null, 1, 10, 32, 255.0, null) == _C.EINVAL      # smaller than the 11x11 window
    assert b"11x11" in L.ofsv_last_error()
    assert L.ofsv_sq_err_f64(null, null, null, null, 0, 16, 1.0, null) == _C.OK
    d = _C.ConvDesc()
    assert L.ofsv_conv_simt(ctypes.byref(d), null, null, null, null, null, null, null) == _C.EINVAL
    with pytest.raises(RuntimeError):
        _C.check(_C.EINVAL)
    assert L.ofsv_launch_count() == n0
    # empty batches are accepted and are no-ops
    assert L.ofsv_warp2d_f32(null, null, null, null, null, 0, 1, 4, 4, 0, null) == _C.OK
    assert L.ofsv_blend_f32(null, null, null, null, 0, null) == _C.OK


def test_interpolation_drivers_mirror_reference_loops():
    """interp.py vs Flow-2D/inference_img.py:64-97 with a stand-in model whose `inference` averages its inputs: the recursive
    driver must return the 2^exp + 1 uniformly spaced members, the ratio driver must bisect to the requested time."""
    from opticalflowscivis_b200 import interp

    class Lerp:
        calls = 0

        def inference(self, a, b):
            Lerp.calls += 1
            return ((a + b) / 2,)

    a, b = torch.zeros(1, 1, 4, 4), torch.ones(1, 1, 4, 4)
    seq = interp.interpolate_recursive(Lerp(), a, b, exp=3)
    assert len(seq) == 9 and Lerp.calls == 7
    assert all(torch.allclose(x, torch.full_like(x, i / 8)) for i, x in enumerate(seq))
    Lerp.calls = 0
    out = interp.interpolate_ratio(Lerp(), a, b, ratio=0.3, rthreshold=0.02, rmaxcycles=8)
    assert len(out) == 3 and abs(float(out[1].mean()) - 0.3) <= 0.01 + 1e-6 and Lerp.calls <= 8
    assert interp.interpolate_ratio(Lerp(), a, b, ratio=0.005)[1] is a and interp.interpolate_ratio(Lerp(), a, b, ratio=0.995)[1] is b


def test_raw_volume_reader_matches_reference_loader(tmp_path):
    """pipeline.read_raw_volume == np.fromfile(...).resize(...) of Datasets/read_data.py:116-119 (x fastest, uint8), incl. the
    zero padding of a short file; raw_pairs yields (i, i + 2) members of the sorted list."""
    import numpy as np
    from opticalflowscivis_b200 import pipeline
    rng = np.random.RandomState(0)
    shape = (6, 8, 10)
    files = []
    for k in range(5):
        v = rng.randint(0, 256, size=shape).astype(np.uint8)
        f = tmp_path / f"drop_{k:04d}.raw"
        v.tofile(f)
        files.append(str(f))
    ref = np.fromfile(files[1], dtype="uint8")
    ref.resize(*shape)
    got = pipeline.read_raw_volume(files[1], shape)
    assert got.shape == (1, 1) + shape and got.dtype == torch.uint8 and np.array_equal(got[0, 0].numpy(), ref)
    short = tmp_path / "short.raw"
    ref[:3].tofile(short)
    ref2 = np.fromfile(short, dtype="uint8")
    ref2.resize(*shape)
    assert np.array_equal(pipeline.read_raw_volume(str(short), shape)[0, 0].numpy(), ref2)
    pairs = list(pipeline.raw_pairs(list(reversed(files)), shape))
    assert len(pairs) == 2
    assert np.array_equal(pairs[1][0][0, 0].numpy(), np.fromfile(files[2], dtype="uint8").reshape(shape))
    assert np.array_equal(pairs[1][1][0, 0].numpy(), np.fromfile(files[4], dtype="uint8").reshape(shape))
    with pytest.raises(ValueError):
        pipeline.read_raw_volume(files[0], shape, out=torch.zeros(5, dtype=torch.uint8))


def test_cpu_tensors_are_rejected():
    from opticalflowscivis_b200 import ops
    from opticalflowscivis_b200.upflow import CorrelationFunction
    with pytest.raises(TypeError):
        ops.warp2d(torch.zeros(1, 1, 4, 4), torch.zeros(1, 2, 4, 4))
    with pytest.raises(TypeError):
        ops.warp3d(torch.zeros(1, 1, 4, 4, 4), torch.zeros(1, 3, 4, 4, 4))
    with pytest.raises(TypeError):
        ops.sq_err_sums(torch.zeros(1, 16), torch.zeros(1, 16))
    with pytest.raises(TypeError):
        ops.ssim2d_means(torch.zeros(16, 16), torch.zeros(16, 16))
    with pytest.raises(TypeError):
        CorrelationFunction.apply(torch.zeros(1, 4, 8, 8), torch.zeros(1, 4, 8, 8), 4, 1, 4, 1, 1, 1)
    with pytest.raises(NotImplementedError):
        CorrelationFunction.apply(torch.zeros(1, 4, 8, 8), torch.zeros(1, 4, 8, 8), 3, 3, 20, 1, 2, 1)


# ------------------------------------------------------------------------------------------------- multi-GPU host logic
def test_shard_range_partitions_every_item_once():
    from opticalflowscivis_b200.shard import shard_range, shard_sizes
    for n in (0, 1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            owned = []
            for r in range(world):
                lo, hi = shard_range(n, r, world)
                assert 0 <= lo <= hi <= n
                owned += list(range(lo, hi))
            assert owned == list(range(n))
            assert max(shard_sizes(n, world)) - min(shard_sizes(n, world)) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _gloo_worker(rank, world, port, n_items, q):
    import torch.distributed as dist
    from opticalflowscivis_b200.shard import gather_counts, shard_range
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(n_items, rank, world)
        total, t = gather_counts(hi - lo, 10.0 + rank)          # rank r "took" 10+r ms
        q.put((rank, lo, hi, total, t))
    finally:
        dist.destroy_process_group()


def test_batch_sharding_world2_gloo():
    """N>1 path on CPU: two gloo ranks shard 5 pairs with no data-path collective; the only reduction is
    (sum of processed items, max of device times), exactly what bench.py does over NCCL."""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, 5, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 3), (3, 5)]
    assert all(r[3] == 5 and r[4] == 11.0 for r in res)


def _gloo_grad_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from opticalflowscivis_b200.optim import GradientBucket, allreduce_gradients
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.PReLU(7), torch.nn.Linear(7, 3))
    bucket = GradientBucket(net.parameters())
    x = torch.full((4, 5), float(rank + 1))
    net(x).sum().backward()                              # autograd accumulates INTO the bucket views
    local = bucket.flat.clone()
    assert all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in net.parameters())
    scale = allreduce_gradients(bucket)
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    ok = torch.allclose(bucket.flat, sum(gathered)) and scale == 1.0 / world and bucket.flat.numel() == sum(p.numel() for p in net.parameters())
    q.put((rank, bool(ok), float(bucket.flat.abs().sum())))
    dist.destroy_process_group()


def test_gradient_bucket_allreduce_world2_gloo():
    """Training-tier collective: the gradients of all parameters are ONE flat bucket, summed by one all_reduce;
    the 1/world average is returned as the optimizer's grad_scale.  Two gloo ranks on CPU."""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res) and res[0][2] == res[1][2] and res[0][2] > 0
    # no process group: no-op, scale 1
    import torch
    from opticalflowscivis_b200.optim import GradientBucket, allreduce_gradients
    b = GradientBucket(torch.nn.Linear(2, 2).parameters())
    assert allreduce_gradients(b) == 1.0


@pytest.mark.parametrize("nd", [2, 3])
def test_stacked_conv_op_lists_cover_every_term_once(nd):
    """csrc/conv_stack.cu builds its MMA list on the host: every (phase, tap, output slice) term must appear exactly once, the
    first MMA into a TMEM column block must be the one that overwrites, and every run must be contiguous in weight rows and
    columns.  ofsv_conv_stack_selfcheck replays the list with the launch code's own plan functions (no GPU needed); the
    modelled tensor cycles per output slice must not get worse with a deeper super-tile (the point of stacking along N)."""
    L = _C.lib()
    for c, cin in ((128, 2), (96 if nd == 2 else 64, 5 + 2 * nd), (64, 5 + 2 * nd)):
        torch.manual_seed(3)
        blk = ifnet.IFBlock(nd, cin, c)
        layers = list(blk.layers()) + [blk._heads_shuffle]
        for li, lay in enumerate(layers):
            if lay.in_stride != 1:
                continue
            d = lay._structure_desc()
            layout = L.ofsv_conv_halo_weight_layout(ctypes.byref(d))
            assert layout in (_C.WL_TAP, _C.WL_STACK)
            if layout != _C.WL_STACK:
                continue
            per_slice = {}
            for td in ((1, 2, 4) if nd == 3 else (1,)):
                nops, cyc = ctypes.c_int(0), ctypes.c_double(0.0)
                rc = L.ofsv_conv_stack_selfcheck(ctypes.byref(d), td, ctypes.byref(nops), ctypes.byref(cyc))
                if rc == _C.ENOSUP:
                    continue
                assert rc == 0, (nd, c, li, td, L.ofsv_last_error().decode())
                assert nops.value >= 1
                per_slice[td] = cyc.value / td
            assert per_slice, (nd, c, li)
            tds = sorted(per_slice)
            for a, b in zip(tds, tds[1:]):
                assert per_slice[b] <= per_slice[a] * 1.001, (nd, c, li, per_slice)
    # the point of the design: a 64 -> 64 3^3 conv at TD = 4 (two issuers, runs of N <= 128 each) needs < 0.85 of the port-bound
    # per-slice MMA time at TD = 1
    blk = ifnet.IFBlock(3, 11, 64)
    d = blk.layers()[2]._structure_desc()
    c1, c4 = ctypes.c_double(0.0), ctypes.c_double(0.0)
    assert L.ofsv_conv_stack_selfcheck(ctypes.byref(d), 1, None, ctypes.byref(c1)) == 0
    assert L.ofsv_conv_stack_selfcheck(ctypes.byref(d), 4, None, ctypes.byref(c4)) == 0
    assert c4.value / 4 < 0.85 * c1.value, (c1.value, c4.value)
