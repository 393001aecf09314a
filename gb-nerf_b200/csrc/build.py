"""Build libgbnerf.so in-tree with plain nvcc for sm_100a (no torch headers, no JIT cache).

    python gb-nerf_b200/csrc/build.py [--force] [--verbose] [--diag]

--diag builds gb-nerf_b200/libgbnerf_diag.so instead: the same library with the diagnostic switches of the MLP
kernels compiled in (-DGBN_TS_DIAG: GBNERF_TS_CHAOS, GBNERF_TS_GATE_DIRECT); select it with GBNERF_LIB=<path>.
--exp builds gb-nerf_b200/libgbnerf_exp.so: the timing experiments of the two-tile MLP kernel (-DGBN_T2_EXP, see
csrc/mlp_t2.cuh); results of its ablations are wrong by design - never load it outside tools/.

The shared library lands next to the package (gb-nerf_b200/libgbnerf.so); it is git-ignored but travels to
the GPU box with the gpurun snapshot.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "libgbnerf.so")
STAMP = os.path.join(PKG, ".libgbnerf.stamp")
SOURCES = ["abi.cu", "render_ops.cu", "mlp_tc.cu", "mlp_aux.cu", "mlp_wgrad.cu", "render_staged.cu", "ray_setup.cu", "tcnn_model.cu", "normals.cu", "mlp_ts.cu"]
# probes that only the diag / exp builds carry (tools/ts_probe*.py bind them with ctypes themselves)
TOOL_SOURCES = [os.path.join(os.path.dirname(PKG), "tools", "experiments", "ts_probe.cu")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC,-O2,-Wall", "--expt-relaxed-constexpr"]


def _digest():
    h = hashlib.sha256(" ".join(FLAGS).encode())
    names = sorted(f for f in os.listdir(HERE) if f.endswith((".cu", ".cuh", ".h")))
    for n in names + ["../../include/gbnerf.h", "../../tools/experiments/ts_probe.cu"]:
        with open(os.path.join(HERE, n), "rb") as fh:
            h.update(n.encode() + b"\0" + fh.read())
    return h.hexdigest()


def build(force=False, verbose=False, diag=False, exp=False, variant=""):
    """variant (with exp): "NAME:MACRO[,MACRO...]" builds libgbnerf_exp_NAME.so with those extra -D flags (A/B of a code
    change on one box in one gpurun call, tools/t2_exp.sh)."""
    srcs = [os.path.join(HERE, s) for s in SOURCES if os.path.exists(os.path.join(HERE, s))]
    if diag or exp:
        srcs += [t for t in TOOL_SOURCES if os.path.exists(t)]
    dig = _digest()
    vname, _, vdefs = variant.partition(":")
    extra = ["-D" + d for d in vdefs.split(",") if d]
    suffix = "_exp" + ("_" + vname if vname else "")
    target = OUT.replace("libgbnerf.so", "libgbnerf_diag.so") if diag else OUT.replace("libgbnerf.so", f"libgbnerf{suffix}.so") if exp else OUT
    stamp = STAMP + (".diag" if diag else "." + suffix[1:] if exp else "")
    dig += "|" + variant
    if not force and os.path.exists(target) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
        return target
    objs = []
    procs = []
    for s in srcs:
        o = os.path.join(HERE, os.path.basename(s).replace(".cu", ".diag.o" if diag else f".{suffix[1:]}.o" if exp else ".o"))
        cmd = [NVCC, *FLAGS, "-I", HERE, *(["-DGBN_TS_DIAG"] if diag else ["-DGBN_T2_EXP", *extra] if exp else []), "-c", s, "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if out.strip() and (verbose or p.returncode != 0):
            print(f"--- {s}\n{out}", file=sys.stderr)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.check_call([NVCC, "-shared", "-o", target, *objs, "-lcudart_static", "-lpthread", "-ldl", "-lrt"])
    for o in objs:
        os.remove(o)
    with open(stamp, "w") as fh:
        fh.write(dig)
    return target


if __name__ == "__main__":
    var = next((a.split("=", 1)[1] for a in sys.argv if a.startswith("--variant=")), "")
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, diag="--diag" in sys.argv, exp="--exp" in sys.argv or bool(var),
                variant=var))
