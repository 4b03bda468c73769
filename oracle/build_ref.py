"""Recipe that makes the UNMODIFIED reference runnable on the GPU box (test / bench infrastructure).

    python oracle/build_ref.py            # in the build container, where /root/reference exists

The reference is pure Python, so "compiling it from the sources where they lie" means byte-compiling the
three modules of the hot path (``py_compile``, CPython 3.12 of this image -- the GPU box runs the same image)
into ``oracle/_ref/*.bin`` (CPython bytecode files; named .bin because snapshot tools commonly drop ``*.pyc``).  No reference SOURCE is copied: ``oracle/_ref/`` holds bytecode only, is listed in
``.gitignore`` (stays out of history) and NOT in ``.gpurunignore`` (travels to the GPU box like the built
``.so``).  ``oracle/ref_verbatim.py`` loads these modules (falling back to the sources under /root/reference
when they are present and ``_ref`` is not) behind two in-process stubs for the absent ``open3d`` / ``opals``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may execute anything under ``oracle/``.
"""
import hashlib
import json
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MORE3D_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")

# module name -> path below the reference root (the §8(a) files of SURVEY.md)
MODULES = {
    "DM_saliency": "DM_saliency.py",                              # Curvature_Kernel / Saliency_Kernel (H5, H6)
    "levelsets_func": "levelsets_func.py",                        # L1-L7
    "BallTreePointSet": "Downsampling/BallTreePointSet.py",       # H1-H4
}


def build(verbose=True):
    if not os.path.isdir(REF):
        if verbose:
            print("oracle/build_ref.py: %s is absent (GPU box): keeping the prebuilt oracle/_ref" % REF)
        return False
    os.makedirs(OUT, exist_ok=True)
    manifest = {"python": sys.version.split()[0], "magic": __import__("importlib.util").util.MAGIC_NUMBER.hex(),
                "modules": {}}
    for name, rel in MODULES.items():
        src = os.path.join(REF, rel)
        dst = os.path.join(OUT, name + ".bin")
        # dfile: the path tracebacks will show -- the reference's own file:line
        py_compile.compile(src, cfile=dst, dfile="reference/" + rel, doraise=True,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
        with open(src, "rb") as f:
            manifest["modules"][name] = {"source": rel, "sha256": hashlib.sha256(f.read()).hexdigest()}
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    if verbose:
        print("oracle/_ref: byte-compiled %s from %s" % (", ".join(sorted(MODULES)), REF))
    return True


if __name__ == "__main__":
    build()
