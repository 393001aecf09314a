"""Recipe that stages the UNMODIFIED reference hot path under ``oracle/_ref/`` (git-ignored, travels with gpurun).

TEST / BENCH INFRASTRUCTURE.  ``/root/reference`` exists only in the build container; ``bench.py --impl reference``
and ``cpu_baseline`` run on the GPU box.  This script (run by ``__graft_entry__.build()`` when the reference tree is
present) copies, byte for byte and into the git-ignored ``oracle/_ref/`` only:

  * ``DS_NeRF/run_nerf_helpers.py`` and ``DS_NeRF/loss.py``  -> ``_ref/DS_NeRF/`` (whole files);
  * from ``run.py`` (which cannot be imported: module-scope CUDA / tkinter / lpips side effects) the
    source text of the hot-path ``FunctionDef``s ``batchify, run_network, batchify_rays, render, render_rays,
    create_nerf`` (run.py:1624-1748, 2003-2128, 2235-2381), verbatim, -> ``_ref/run_hotpath.py``;
  * ``MANIFEST.json``: sha256 of each source file and the line ranges taken.

Nothing staged here is imported by the product package; ``oracle/ref_loader.load_staged()`` is the only consumer.
Reference sources are never committed: ``oracle/_ref/`` is listed in ``.gitignore``.

    python oracle/make_ref.py [--reference /root/reference]
"""
import ast
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
WANTED = ("batchify", "run_network", "batchify_rays", "render", "render_rays", "create_nerf")


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def make(reference="/root/reference"):
    run_py = os.path.join(reference, "run.py")
    if not os.path.isfile(run_py):
        return None
    os.makedirs(os.path.join(OUT, "DS_NeRF"), exist_ok=True)
    manifest = {"reference": reference, "files": {}, "functions": {}}
    for rel in ("DS_NeRF/run_nerf_helpers.py", "DS_NeRF/loss.py"):
        src = os.path.join(reference, rel)
        shutil.copyfile(src, os.path.join(OUT, rel))
        manifest["files"][rel] = sha(src)
    open(os.path.join(OUT, "DS_NeRF", "__init__.py"), "w").close()
    text = open(run_py).read()
    manifest["files"]["run.py"] = sha(run_py)
    parts = ["# Verbatim FunctionDefs of the reference run.py (staged by oracle/make_ref.py; do not edit, do not commit)\n"]
    for node in ast.parse(text).body:
        if isinstance(node, ast.FunctionDef) and node.name in WANTED:
            parts.append(ast.get_source_segment(text, node) + "\n")
            manifest["functions"][node.name] = [node.lineno, node.end_lineno]
    missing = [w for w in WANTED if w not in manifest["functions"]]
    if missing:
        raise RuntimeError(f"reference run.py lacks {missing}")
    with open(os.path.join(OUT, "run_hotpath.py"), "w") as fh:
        fh.write("\n\n".join(parts))
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as fh:
        json.dump(manifest, fh, indent=1)
    return OUT


if __name__ == "__main__":
    ref = sys.argv[sys.argv.index("--reference") + 1] if "--reference" in sys.argv else "/root/reference"
    print(make(ref) or f"no reference tree at {ref}: nothing staged")
