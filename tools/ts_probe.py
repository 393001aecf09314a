import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
# the probe lives in the exp / diag builds only (csrc/build.py --exp): bound here, not in the product's _lib table
_dll = C.CDLL(os.environ.get("GBNERF_LIB") or os.path.join(ROOT, "gb-nerf_b200", "libgbnerf_exp.so"))
class _lib:
    @staticmethod
    def call(name, *args):
        rc = getattr(_dll, name)(*args)
        assert rc == 0, (name, rc)
torch.manual_seed(0)
A = torch.randn(128, 64).bfloat16().cuda()
B = torch.randn(128, 64).bfloat16().cuda()
# K-major SW128 image: row n, 16-byte chunk c at position c ^ (n & 7)
img = torch.empty(128, 8, 8, dtype=torch.bfloat16, device="cuda")
n = torch.arange(128, device="cuda")[:, None]; c = torch.arange(8, device="cuda")[None, :]
img[n.expand(128, 8), (c ^ (n & 7))] = B.view(128, 8, 8)
want = A.float() @ B.float().t()
for a_col, cpk in ((256, 8), (0 + 128, 8)):
    D = torch.zeros(128, 128, device="cuda")
    _lib.call("gbn_debug_ts_mma", C.c_void_p(A.data_ptr()), C.c_void_p(img.data_ptr()), C.c_void_p(D.data_ptr()), a_col, cpk,
              C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    print(f"a_col {a_col} cols/kstep {cpk}: max abs err {(D - want).abs().max().item():.4e} (|want| max {want.abs().max().item():.2f})")
