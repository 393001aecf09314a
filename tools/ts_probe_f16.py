"""How does tcgen05.mma lay out an fp16 accumulator in TMEM?  D = A[128,64] B[128,64]^T with fp16 operands and fp16 D."""
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
A = (torch.randn(128, 64) * 0.5).half().cuda()
B = (torch.randn(128, 64) * 0.5).half().cuda()
img = torch.empty(128, 8, 8, dtype=torch.float16, device="cuda")
n = torch.arange(128, device="cuda")[:, None]; c = torch.arange(8, device="cuda")[None, :]
img[n.expand(128, 8), (c ^ (n & 7))] = B.view(128, 8, 8)
want = (A.float() @ B.float().t())
raw = torch.zeros(128, 128, dtype=torch.int32, device="cuda")
_lib.call("gbn_debug_ts_mma_f16", C.c_void_p(A.data_ptr()), C.c_void_p(img.data_ptr()), C.c_void_p(raw.data_ptr()), 256, 8,
          C.c_void_p(torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
cells = raw.cpu()
lo = (cells & 0xffff).to(torch.int16).view(torch.float16).float()
hi = ((cells >> 16) & 0xffff).to(torch.int16).view(torch.float16).float()
w = want.cpu()
print("nonzero cells per row (first rows):", [(cells[r] != 0).sum().item() for r in range(4)])
# hypothesis 1: packed, column c holds (n = 2c, 2c+1)
h1 = torch.stack([lo[:, :64], hi[:, :64]], -1).reshape(128, 128)
print("packed (2c, 2c+1) in cols 0..63:  max err", (h1 - w).abs().max().item())
# hypothesis 2: one fp16 per column (low half)
print("one per column (low half):        max err", (lo - w).abs().max().item())
# hypothesis 3: column c holds (n = c, c + 64)
h3 = torch.cat([lo[:, :64], hi[:, :64]], -1)
print("packed (c, c+64):                 max err", (h3 - w).abs().max().item())
print("reference magnitude", w.abs().max().item(), " fp16-accumulate rounding ~", w.abs().max().item() * 2 ** -10)
