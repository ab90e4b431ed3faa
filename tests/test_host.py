"""CPU-side contract tests: the C-ABI library loads and exports what include/pcv_b200.h declares, get_model keeps
the reference's contract, the plan compiler lowers every config and places buffers without overlap."""
import ctypes
import os
import re

import pytest
import torch
import torch.nn as nn

import pytorchcv_b200 as P
from pytorchcv_b200 import _lib, blocks as B, nets as M, plan as PL
from pytorchcv_b200._lib import BF16, F32
from conftest import ROOT


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "pcv_b200.h")).read()
    return sorted(set(re.findall(r"PCV_API[^;(]*?\b(pcv_[a-z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_every_header_symbol():
    lib = _lib.load()
    names = _header_symbols()
    assert len(names) >= 24
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pcv_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes SIGNATURES out of sync with the header"
    assert lib.pcv_version() >= 100


def test_no_device_is_an_error_not_a_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    arch = ctypes.c_int()
    rc = _lib.load().pcv_device_info(0, ctypes.byref(arch), None, None)
    assert rc == _lib.ERR_NO_DEVICE
    assert b"no CPU fallback" in _lib.load().pcv_last_error()
    net = P.get_model("resnet18").eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.zeros(1, 3, 224, 224))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        B.conv3x3_block(in_channels=8, out_channels=8).eval()(torch.zeros(1, 8, 4, 4))


def test_abi_validates_arguments():
    d = _lib.ConvDesc(N=1, H=8, W=8, Cin=6, Cout=8, kh=3, kw=3, stride=1, pad=1, dil=1, groups=4, act=0)
    wb, bb = ctypes.c_size_t(), ctypes.c_size_t()
    rc = _lib.load().pcv_conv_packed_bytes(ctypes.byref(d), BF16, ctypes.byref(wb), ctypes.byref(bb))
    assert rc == _lib.ERR_INVALID and b"groups" in _lib.load().pcv_last_error()
    d.groups, d.act = 1, 99
    assert _lib.load().pcv_conv_packed_bytes(ctypes.byref(d), BF16, ctypes.byref(wb), ctypes.byref(bb)) == _lib.ERR_INVALID
    with pytest.raises(_lib.PcvError):
        _lib.call("pcv_plan_run", None, None)


def test_packed_sizes_follow_route():
    wb, bb = ctypes.c_size_t(), ctypes.c_size_t()
    d = _lib.ConvDesc(N=1, H=14, W=14, Cin=256, Cout=256, kh=3, kw=3, stride=1, pad=1, dil=1, groups=1, act=1)
    _lib.call("pcv_conv_packed_bytes", ctypes.byref(d), BF16, ctypes.byref(wb), ctypes.byref(bb))
    assert wb.value == 256 * 9 * 256 * 2 and bb.value == 256 * 4          # tcgen05 route: bf16 [Cout, taps*Cpad]
    _lib.call("pcv_conv_packed_bytes", ctypes.byref(d), F32, ctypes.byref(wb), ctypes.byref(bb))
    assert wb.value == 256 * 9 * 256 * 4                                    # fp32 tier: CUDA-core route
    d.groups = 256
    _lib.call("pcv_conv_packed_bytes", ctypes.byref(d), BF16, ctypes.byref(wb), ctypes.byref(bb))
    assert wb.value == 256 * 9 * 4                                          # depthwise: fp32 [tap][C]


def test_dual_source_conv_domain():
    """pcv_conv1x1_dual_ok (host logic only): ResNet-50's four projection units are inside the dual-source kernel's domain in
    both 16-bit tiers; narrow outputs, the fp32 tier, mismatched grids and an activated shortcut conv are not."""
    from pytorchcv_b200._lib import F16
    lib = _lib.load()

    def pair(cin, mid, cout, stride, hw, n=256, **kw2):
        ho = (hw - 1) // stride + 1
        d = _lib.ConvDesc(N=n, H=ho, W=ho, Cin=mid, Cout=cout, kh=1, kw=1, stride=1, pad=0, dil=1, groups=1, act=_lib.ACT_RELU)
        f2 = dict(N=n, H=hw, W=hw, Cin=cin, Cout=cout, kh=1, kw=1, stride=stride, pad=0, dil=1, groups=1, act=_lib.ACT_NONE)
        f2.update(kw2)
        return d, _lib.ConvDesc(**f2)

    for cin, mid, cout, stride, hw in [(64, 64, 256, 1, 56), (256, 128, 512, 2, 56), (512, 256, 1024, 2, 28), (1024, 512, 2048, 2, 14)]:
        d, d2 = pair(cin, mid, cout, stride, hw)
        for tier in (BF16, F16):
            assert lib.pcv_conv1x1_dual_ok(ctypes.byref(d), ctypes.byref(d2), tier) == 1, (cin, cout, tier)
        assert lib.pcv_conv1x1_dual_ok(ctypes.byref(d), ctypes.byref(d2), F32) == 0
        d.flags = _lib.CONV_SE_GATE   # the gated variant (SE units) shares the domain
        assert lib.pcv_conv1x1_dual_ok(ctypes.byref(d), ctypes.byref(d2), BF16) == 1
    for bad in (pair(64, 32, 128, 1, 56),                       # Cout <= 128: not the 256-wide pair tile
                pair(256, 128, 512, 2, 56, act=_lib.ACT_RELU),  # the shortcut conv must be linear
                pair(256, 128, 512, 2, 56, N=128),              # different batch
                pair(256, 128, 512, 2, 56, H=54, W=54),         # grids do not meet
                pair(256, 128, 512, 2, 56, Cout=256),           # widths do not meet
                pair(1024, 512, 2048, 2, 14, n=2)):             # a single 128-row tile: outside the CTA-pair kernel
        assert lib.pcv_conv1x1_dual_ok(ctypes.byref(bad[0]), ctypes.byref(bad[1]), BF16) == 0
    assert lib.pcv_conv1x1_dual_ok(None, None, BF16) == 0


def test_get_model_contract():
    assert isinstance(P.get_model("ResNet18"), nn.Module)                   # case-insensitive (model_provider.py:1378)
    with pytest.raises(ValueError, match="Unsupported model"):
        P.get_model("resnet19")
    with pytest.raises(ValueError, match="Unsupported ResNet"):
        M.get_resnet(blocks=19)
    with pytest.raises(ValueError, match="model_name"):
        M.get_resnet(blocks=18, pretrained=True)
    with pytest.raises(NotImplementedError):
        B.create_activation_layer("gelu")                                   # activ.py:219
    for n in ("resnet18", "resnet50", "mobilenetv2_w1", "seresnext50_32x4d", "deeplabv3_resnetd50b_voc"):
        assert n in P.supported_models()


@pytest.mark.parametrize("name,count", [("resnet18", 11689512), ("resnet50", 25557032), ("resnet50b", 25557032),
                                        ("mobilenetv2_w1", 3504960), ("seresnext50_32x4d", 27559896),
                                        ("resnext50_32x4d", 25028904), ("mobilenet_w1", 4231976),
                                        ("deeplabv3_resnetd50b_voc", 42127850)])
def test_parameter_counts_pinned_by_reference_tests(name, count):
    """resnet.py:975-995, mobilenetv2.py:436, seresnext.py:296, resnext.py, mobilenet.py, deeplabv3.py:678."""
    net = P.get_model(name, pretrained=False)
    assert sum(p.numel() for p in net.parameters() if p.requires_grad) == count


def test_kwargs_flow_to_constructor():
    net = P.get_model("resnet18", in_channels=1, num_classes=10, in_size=(64, 64))
    assert net.features.init_block.conv.conv.in_channels == 1 and net.output.out_features == 10
    dl = P.get_model("deeplabv3_resnetd50b_voc", aux=False, num_classes=5)
    assert not hasattr(dl, "aux_block") and dl.final_block.conv2.out_channels == 5


CONFIGS = [("resnet18", (8, 3, 224, 224), 23), ("resnet50", (4, 3, 224, 224), 56), ("mobilenetv2_w1", (4, 3, 224, 224), 55),
           ("seresnext50_32x4d", (4, 3, 224, 224), 104), ("deeplabv3_resnetd50b_voc", (1, 3, 480, 480), 70),
           ("efficientnet_b0", (4, 3, 224, 224), 99), ("mobilenetv3_large_w1", (4, 3, 224, 224), 73)]


@pytest.mark.parametrize("tier", [BF16, F32])
@pytest.mark.parametrize("name,shape,n_ops", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_lowering_and_arena(name, shape, n_ops, tier):
    """Dry-run the compiler on CPU: op counts, output shapes, and no two live buffers share arena bytes."""
    net = P.get_model(name).eval()
    b = PL.Builder(tier, torch.device("cpu"))
    N, _, H, W = shape
    x = b.new(N, H, W, 8)
    x.buf.first, x.buf.pinned = -1, True
    out = PL.lower(b, net, x)
    _, trefs = PL._flatten(out)
    # plan ops + the per-call edge ops (fp32 NCHW outputs written after the plan into fresh, caller-owned tensors);
    # on the 16-bit tiers 13 of MobileNetV2's 17 units (>= 14x14 outputs) have their dw -> pw pair fused (-1 op each) and the 3
    # stride-2 ones among them are ONE fused expansion -> dw -> pw op (-1 more each)
    n_fused = 16 if (name == "mobilenetv2_w1" and tier == BF16) else 0
    if name == "seresnext50_32x4d" and tier == BF16:
        # the SE scale + identity + ReLU of all 16 units rides on conv3's epilogue (PCV_CONV_SE_GATE); the 4 projection shortcuts
        # are the second half of that conv's K dimension (pcv_conv1x1_dual_se)
        n_fused = 16 + 2   # (the gated variant is taken for output maps of >= 28 x 28: stages 1 and 2)
    if name in ("resnet50", "deeplabv3_resnetd50b_voc") and tier == BF16:
        n_fused = 4    # the 4 projection shortcuts are K-concatenated into their units' conv3 (pcv_conv1x1_dual)
    assert len(b.ops) + sum(t.tail is not None for t in trefs) == n_ops - n_fused
    for t in trefs:
        t.buf.pinned = True
    if name.startswith("deeplab"):
        assert [(t.N, t.C, t.H, t.W, t.layout) for t in trefs] == [(N, 21, 480, 480, "nchw")] * 2
        assert all(t.tail is not None and t.buf.nbytes == 0 for t in trefs)   # no arena storage: nothing to alias
    else:
        assert [(t.N, t.C, t.flat) for t in trefs] == [(N, 1000, True)]
    top = PL._assign_offsets(b.bufs)
    assert top < sum(z.nbytes for z in b.bufs)  # liveness reuse actually happens
    life = lambda z: (z.first, 1 << 60 if z.pinned else max(z.last, z.first))
    bufs = b.bufs
    for i in range(len(bufs)):
        for j in range(i + 1, len(bufs)):
            a, c = bufs[i], bufs[j]
            (a0, a1), (c0, c1) = life(a), life(c)
            if a1 < c0 or c1 < a0:
                continue
            assert a.offset + a.nbytes <= c.offset or c.offset + c.nbytes <= a.offset, "live buffers overlap"


def test_identity_buffer_outlives_the_unit():
    """SURVEY hard part 8: the unit's input must stay intact until the last conv's epilogue has read it."""
    unit = M.ResUnit(64, 64, stride=1, bottleneck=True).eval()
    b = PL.Builder(BF16, torch.device("cpu"))
    x = b.new(2, 14, 14, 64)
    PL.lower(b, unit, x)
    assert x.buf.last == len(b.ops) - 1 == 2  # read by conv1 (op 0) and again as the residual of conv3 (op 2)


def test_unsupported_patterns_raise_at_compile_time():
    b = PL.Builder(BF16, torch.device("cpu"))
    x = b.new(1, 8, 8, 16)
    with pytest.raises(NotImplementedError, match="outside the B200 eval path"):
        PL.lower(b, nn.GELU(), x)
    with pytest.raises(RuntimeError, match="eval-mode"):
        PL.lower(b, B.conv1x1_block(in_channels=16, out_channels=16), x)            # still in training mode
    with pytest.raises(NotImplementedError, match="cannot be folded"):
        PL.lower(b, B.ConvBlock(16, 16, 1, normalization=lambda num_features: nn.InstanceNorm2d(num_features)).eval(), x)
    # a 4-tuple padding (nn.ZeroPad2d in front of the conv, conv.py:245-249) lowers to one zero-pad pass + the conv
    n0 = len(b.ops)
    y = PL.lower(b, B.ConvBlock(16, 16, 3, padding=(1, 0, 1, 0)).eval(), x)
    assert len(b.ops) == n0 + 2 and (y.H, y.W) == (7, 7)
    n0 = len(b.ops)
    y = PL.lower(b, B.ConvBlock(16, 16, 3, padding=(1, 1, 1, 1)).eval(), x)   # symmetric amounts ride on the kernel's padding
    assert len(b.ops) == n0 + 1 and (y.H, y.W) == (8, 8)
    with pytest.raises(RuntimeError, match="eval-mode"):
        P.accelerate(P.get_model("resnet18"))


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "pytorchcv_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f"{f} imports the oracle"
                assert "/root/reference" not in text, f"{f} reads the reference at run time"


def test_pretrained_loads_a_verified_local_checkpoint(tmp_path):
    """pretrained=True: the reference's cache file naming `{model}-{error}-{sha1[:8]}.pth`
    (model_store.py:158-165), content hash checked, never a download."""
    import hashlib
    src = P.get_model("resnet10", pretrained=False)
    with torch.no_grad():
        for prm in src.parameters():
            prm.add_(0.25)
    raw = tmp_path / "w.pth"
    torch.save(src.state_dict(), raw)
    sha = hashlib.sha1(raw.read_bytes()).hexdigest()
    raw.rename(tmp_path / f"resnet10-1234-{sha[:8]}.pth")
    net = P.get_model("resnet10", pretrained=True, root=str(tmp_path))
    for (ka, a), (kb, bb) in zip(src.state_dict().items(), net.state_dict().items()):
        assert ka == kb and torch.equal(a, bb)
    # a file whose content does not match the hash in its name is rejected; a missing file is an error, not a download
    (tmp_path / "resnet12-0000-deadbeef.pth").write_bytes(b"not a checkpoint")
    with pytest.raises(RuntimeError, match="sha1 mismatch"):
        P.get_model("resnet12", pretrained=True, root=str(tmp_path))
    with pytest.raises(RuntimeError, match="never downloads"):
        P.get_model("resnet14", pretrained=True, root=str(tmp_path))
