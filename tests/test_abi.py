"""CPU: the C-ABI library builds, loads and exports every symbol the header declares; the drop-in
modules expose the reference's key layout and refuse to run without CUDA (no fallback)."""
import ctypes
import os

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as ge
    ge.build()
    import sisr_b200
    return sisr_b200


def test_library_exports_every_declared_symbol(built):
    from sisr_b200 import _lib
    protos = _lib.parse_header()
    assert len(protos) >= 30
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in protos:
        assert hasattr(lib, name), f"{name} declared in include/sisr_b200.h but not exported"
    lib.sisr_abi_version.restype = ctypes.c_int
    assert lib.sisr_abi_version() == 2


def test_tensor_core_shape_predicate(built):
    from sisr_b200 import _lib
    lib = _lib.load()
    trunk = _lib.ConvDesc(64, 24, 24, 64, 24, 24, 64, 3, 1, 1, 0)
    first = _lib.ConvDesc(64, 24, 24, 3, 24, 24, 64, 9, 1, 4, 0)
    bad = _lib.ConvDesc(64, 24, 24, 64, 99, 24, 64, 3, 1, 1, 0)
    assert lib.sisr_conv_uses_tensor_cores(ctypes.addressof(trunk)) == 1
    assert lib.sisr_conv_uses_tensor_cores(ctypes.addressof(first)) == 0
    assert lib.sisr_conv_uses_tensor_cores(ctypes.addressof(bad)) == 0


def test_state_dict_layout_matches_reference_keys(built):
    from oracle import state_factory as S
    m = built
    g = m.GeneratorSuffix(m.GeneratorSuffix(m.Generator(16, 64, 256, [2], use_sn=True)))
    st = S.generator_state(0, n_blocks=16, n_suffix=2)
    assert set(g.state_dict().keys()) == set(st.keys())
    assert len(m.Generator(16, 64, 256, [2], use_sn=True).state_dict()) == 327
    for k, v in g.state_dict().items():
        assert tuple(v.shape) == tuple(st[k].shape), k
    d = m.Discriminator((3, 96, 96), [64, 64, 128, 128, 256, 256, 512, 512], [1, 2, 1, 2, 1, 2, 1, 2])
    assert d.fc_in == 18432 and d.fc_mid == 1024
    assert len(d.state_dict()) == 71
    assert sum(p.numel() for p in d.parameters()) == 23565505
    assert set(d.state_dict().keys()) == set(S.discriminator_state(0).keys())
    v = m.MaskedVGG(0b10000)
    assert len(v.state_dict()) == 32 and len(v.layers) == 35
    assert not any(p.requires_grad for p in v.parameters())
    no_sn = m.Generator(2, 64, 256, [2], use_sn=False).state_dict()
    assert "upscale.0.0.weight" in no_sn and "end.0.weight" in no_sn
    assert "first_layers.0.weight_orig" in no_sn      # trunk is spectral-normed unconditionally


def test_freeze_semantics(built):
    m = built
    g1 = m.Generator(2, 64, 256, [2], use_sn=True)
    g2 = m.GeneratorSuffix(g1, freeze_prefix=True, freeze_upscale=True, freeze_end=True)
    trainable = sorted(k for k, p in g2.named_parameters() if p.requires_grad)
    assert trainable == ["upscale.0.bias", "upscale.0.weight_orig", "upscale.2.weight"]
    assert sum(p.numel() for p in g2.parameters() if p.requires_grad) == 147713
    assert isinstance(g2.end, list) and g2.end[0] is g1.end


def test_generator_x2_to_x4_weight_coverage(built):
    m = built
    x2 = m.Generator(16, 64, 256, [2], use_sn=True)
    x4 = m.GeneratorSuffix(m.Generator(16, 64, 256, [2], use_sn=True))
    n_x2 = sum(p.numel() for p in x2.parameters())
    n_x4 = sum(p.numel() for p in x4.parameters())
    assert (n_x2, n_x4) == (1387925, 1535638)          # README.md:69 "90 %"


def test_no_cpu_fallback(built):
    m = built
    g = m.Generator(1, 64, 256, [2], use_sn=True)
    with pytest.raises(Exception) as e:
        g(torch.zeros(1, 3, 8, 8))
    assert "CUDA" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "single-image-super-resolution_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower().replace("# oracle", ""), f


def test_bench_reads_dominant_kernel_share_from_committed_profile():
    """bench.py's roofline.share_of_step is parsed from the newest committed launch-list summary under profiles/."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("_bench_mod", os.path.join(root, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    text = mod.dominant_share()
    assert text.startswith(("igemm_t", "igemm_pm")) and "%" in text and "launches" in text and "profiles/" in text


def test_bench_configs_and_traffic_source():
    """bench.py: the three BASELINE.json configurations carry SURVEY 8(d)'s FLOP budgets, roofline.traffic is
    read from a committed ncu summary (never a literal), and the CPU arm prefers the real reference."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    assert abs(b.CONFIGS["x4"]["flop"] - 42.55e9) < 1e6 and abs(b.CONFIGS["frozen"]["flop"] - 38.66e9) < 1e6
    assert abs(b.CONFIGS["x8"]["flop"] - 280.75e9) < 1e6 and b.CONFIGS["x8"]["hr"] == 256
    t = b.ncu_traffic()
    # either a committed capture of the kernel the trunk runs NOW (igemm_pm_kernel) or null - never a number measured
    # on a kernel that is no longer on the path
    if t["source"] is None:
        assert t["bytes"] is None
    else:
        assert t["source"].startswith("profiles/") and t["bytes"] > 1e6
        assert os.path.exists(os.path.join(ROOT, t["source"]))
        assert "igemm_pm_kernel" in open(os.path.join(ROOT, t["source"])).read()
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert '"traffic": 4831488' not in src
    assert "reference_train_loop_time" in src and 'kind": "reference"' in src.replace("'", '"')
