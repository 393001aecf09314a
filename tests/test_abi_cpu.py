"""CPU: the C-ABI library builds, loads and exports exactly what include/gbnerf.h declares; host-side logic that
needs no GPU (layout plan, argument validation, the no-fallback rule)."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gbnerf.h")


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge._load_builder().build()
    from gbnerf_b200 import _lib
    return _lib


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gbn_\w+)\s*\(", src)))


def test_header_and_binding_agree(lib):
    names = header_functions()
    assert len(names) >= 15
    assert sorted(lib.SIGNATURES) == names
    so = ctypes.CDLL(lib.LIB_PATH)
    for n in names:
        assert hasattr(so, n), f"{n} declared in gbnerf.h but not exported"


def test_exports_are_plain_c(lib):
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib.LIB_PATH], text=True)
    exported = [l.split()[-1] for l in out.splitlines() if " T " in l]
    gbn = [e for e in exported if e.startswith("gbn_")]
    assert sorted(gbn) == header_functions()


def test_version_and_sizes(lib):
    l = lib.load()
    assert l.gbn_version() == 100
    bf16, tf32 = l.gbn_mlp_packed_bytes(0), l.gbn_mlp_packed_bytes(1)
    # 593,408 weights padded to the UMMA tiling: K 63->64, 319->64+256, alpha/rgb N->16, views K 283->256 (+fp32 dir part)
    assert 1_190_000 < bf16 < 1_300_000 and 2_380_000 < tf32 < 2_600_000
    assert bf16 % 256 == 0 and tf32 % 256 == 0
    assert l.gbn_mlp_packed_bytes(7) == 0
    assert l.gbn_mlp_workspace_bytes(10) == 256 + 10 * 128 * 4


def test_sm100a_tensor_core_sass(lib):
    """The MLP kernel must be a tcgen05 kernel: UTC*MMA + LDTM + UBLKCP in the SASS, no legacy HMMA."""
    sass = subprocess.check_output(["cuobjdump", "-sass", lib.LIB_PATH], text=True)
    assert "sm_100a" in sass
    assert re.search(r"\bUTCHMMA\b", sass), "no tcgen05.mma (UTCHMMA) in SASS"
    assert re.search(r"\bLDTM\b", sass), "no tcgen05.ld (LDTM) in SASS"
    assert re.search(r"\bUBLKCP\b", sass), "no bulk TMA (UBLKCP) in SASS"
    # warp-level HMMA is allowed only in the hash-grid model's kernels (64-wide MLPs under a gather-bound kernel,
    # csrc/tcnn_model.cu); the 8x256 network must not fall back to it
    for fn in re.split(r"\n\s*Function : ", sass)[1:]:
        name = fn.split("\n", 1)[0]
        if re.search(r"\bHMMA\b", fn):
            assert "tcnn" in name, f"legacy HMMA in {name}"
        if "nerf_mlp" in name and "prepack" not in name:
            assert re.search(r"\bUTC[A-Z]*MMA\b", fn), f"{name} has no tcgen05.mma"


def test_experiments_stay_out_of_the_product_build(lib):
    """Timing experiments and probes live in the exp / diag builds only (csrc/build.py --exp): the product library has the
    two shipped schedules of the two-tile MLP kernel (+ its stash-writing form), not the staggered mode, and no probe
    entry points."""
    syms = subprocess.check_output(["nm", lib.LIB_PATH], text=True)
    t2 = sorted(set(re.findall(r"nerf_mlp_t2_kernelILi(\d)ELb(\d)E", syms)))
    assert t2 == [("0", "0"), ("1", "0"), ("1", "1")], t2
    assert "gbn_debug_ts_mma" not in syms and "ts_probe_kernel" not in syms
    assert "mlp_tq" not in syms
    src = open(os.path.join(ROOT, "gb-nerf_b200", "csrc", "build.py")).read()
    sources = re.search(r"^SOURCES = \[(.*?)\]", src, flags=re.M).group(1)
    assert "ts_probe" not in sources and "mlp_tq" not in sources


def test_no_cpu_fallback(lib):
    """Operators refuse CPU tensors; without a GPU compute entry points report a CUDA error, never a result."""
    from gbnerf_b200 import ops
    with pytest.raises(ValueError):
        ops.composite(torch.zeros(2, 4, 4), torch.zeros(2, 4), torch.zeros(2, 3))
    with pytest.raises(ValueError):
        ops.sample_pdf(torch.zeros(2, 5), torch.zeros(2, 4), 8)
    if not torch.cuda.is_available():
        buf = (ctypes.c_float * 64)()
        p = ctypes.cast(buf, ctypes.c_void_p)
        with pytest.raises(lib.GbnError):
            lib.call("gbn_zvals_stratified", p, p, 1, 4, 4, 0, None, p, None)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "gb-nerf_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "nerf_oracle" not in src, f


def test_argument_validation(lib):
    l = lib.load()
    assert l.gbn_zvals_stratified(None, None, 1, 4, 4, 0, None, None, None) == 1   # GBN_EINVAL
    assert b"null" in l.gbn_last_error_string()
    assert l.gbn_sample_pdf_merge(None, None, None, 4, 2, 4, None, None, None, None) == 1
    assert l.gbn_mlp_forward(None, 5, None, None, None, 3, None, None, 4, 4, None, None, None, None) == 1
    assert l.gbn_mlp_backward_data(None, None, 4, None, None, None, None) == 1
    assert l.gbn_mlp_stash_bytes(129) == 2 * 40 * 16384 and l.gbn_mlp_stash_bytes(0) == 0
    assert l.gbn_mlp_packed_bytes(2) > 1_000_000    # transposed bf16 weights for the dgrad pass
    assert l.gbn_pack_rays(None, 0, None, 0, None, 0, None, 0, None, 4, 4, 1.0, 0, 0, 4, 4, 1, 0, 0.0, 1.0, 16, None, None) == 1
    assert l.gbn_adam_step_repack(None, None, None, None, 1e-3, 0.9, 0.999, 1e-8, 1, None, None, None) == 1
    assert l.gbn_mlp_variant() in (0, 1, 2) and l.gbn_kernel_launches() >= 0
    assert l.gbn_composite_forward(None, None, None, 3, None, 0, 64, 1, None, None, None, None, None, None, None) == 0


def test_watchdog_record_is_empty_before_any_launch(lib):
    """gbn_watchdog_report makes no CUDA call: on a machine without a GPU it reports that nothing was recorded, and the
    per-launch check of ops.py passes."""
    assert lib.watchdog_report() is None
    from gbnerf_b200 import ops
    ops._raise_if_watchdog_fired()


def test_proxy_fence_precedes_every_release_of_a_bulk_copied_buffer_read_with_ordinary_loads():
    """Regression guard for the round-1 nondeterminism: a shared-memory buffer filled by a bulk copy (async
    proxy) and read with ordinary loads (generic proxy) must see `fence.proxy.async` before the mbarrier arrival that hands
    it back to the bulk-copy producer - an mbarrier hand-over alone does not order generic reads before async writes.
    The barrier protocol model cannot see this (its phases are correct either way), so the two places are pinned in the
    source: the gate staging of the dgrad epilogue and the G-tile stages read by wgrad's bias warps."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ts = open(os.path.join(root, "gb-nerf_b200", "csrc", "mlp_ts.cu")).read()
    i = ts.index("mbar_arrive(base + L::m_empty + 8 * b);")
    before = ts[ts.rindex("ld_smem16(hb", 0, i):i]
    assert "fence_proxy_async_smem();" in before, "dgrad epilogue: no proxy fence between the gate reads and the m_empty arrival"
    wg = open(os.path.join(root, "gb-nerf_b200", "csrc", "mlp_wgrad.cu")).read()
    j = wg.index("const uint4 v = *reinterpret_cast<const uint4*>(cp + (lane + 32 * h) * 16);")
    k = wg.index("mbar_arrive(base + L::empty + 8 * s);", j)
    assert "fence_proxy_async_smem();" in wg[j:k], "wgrad bias warps: no proxy fence between the tile reads and the stage release"
    # the only other generic reads of bulk-copied shared memory would be new code: flag any ld.shared helper use in the
    # TS kernels outside the gate staging
    assert len(re.findall(r"ld_smem16\(", ts)) == 2, "a new ld_smem16 reader appeared: check its release path (DESIGN 3.2)"


def test_reference_staging_recipe(tmp_path):
    """oracle/make_ref.py copies the reference hot path byte for byte (container only: needs /root/reference)."""
    import hashlib
    import importlib.util
    import json
    import pytest
    if not os.path.isfile("/root/reference/run.py"):
        pytest.skip("reference tree not present (GPU box)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("make_ref", os.path.join(root, "oracle", "make_ref.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.OUT = str(tmp_path / "_ref")
    out = mod.make("/root/reference")
    man = json.load(open(os.path.join(out, "MANIFEST.json")))
    for rel in ("DS_NeRF/run_nerf_helpers.py", "DS_NeRF/loss.py"):
        assert hashlib.sha256(open(os.path.join(out, rel), "rb").read()).hexdigest() == man["files"][rel] == \
            hashlib.sha256(open(os.path.join("/root/reference", rel), "rb").read()).hexdigest()
    src = open("/root/reference/run.py").read().split("\n")
    staged = open(os.path.join(out, "run_hotpath.py")).read()
    for name, (lo, hi) in man["functions"].items():
        assert "\n".join(src[lo - 1:hi]) in staged, name
    assert set(man["functions"]) == {"batchify", "run_network", "batchify_rays", "render", "render_rays", "create_nerf"}
