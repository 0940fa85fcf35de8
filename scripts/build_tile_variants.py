"""Build libfofb200 variants that differ in the compile-time constants of the cell tiles (csrc/tile.cuh, catalog.cuh).
Developer tool for sweeps on the GPU box: FOFB_LIB=<path> makes fof_b200._lib load a variant (scripts/gpu_tile_sweep.sh).
Usage: python scripts/build_tile_variants.py"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fof_b200 import build as B  # noqa: E402


def flags(w=4, threads=128, maxpts=1536, resident=7, zbins=64):
    return [f"-DFOFB_CT_W={w}", f"-DFOFB_CT_THREADS={threads}", f"-DFOFB_CT_MAXPTS={maxpts}", f"-DFOFB_CT_RESIDENT={resident}",
            f"-DFOFB_ZBINS={zbins}"]


VARIANTS = {
    "t256m3072r4": flags(threads=256, maxpts=3072, resident=4),   # the round-2 first build
    "t128m1536r7": flags(),                                       # the default
    "t128m1536r7zb128": flags(zbins=128),
    "w8t256m2432r4": flags(w=8, threads=256, maxpts=2432, resident=4),
}


def main():
    B.build()
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    out = os.path.join(ROOT, "fof_b200", "variants")
    os.makedirs(out, exist_ok=True)
    objdir = os.path.join(ROOT, "fof_b200", "build")
    for name, fl in VARIANTS.items():
        vdir = os.path.join(objdir, "var_" + name)
        os.makedirs(vdir, exist_ok=True)

        def one(src):
            o = os.path.join(vdir, src.replace(".cu", ".o"))
            r = subprocess.run([nvcc] + B.NVCC_FLAGS + fl + ["-c", os.path.join(B.CSRC, src), "-o", o], capture_output=True, text=True)
            if r.returncode:
                sys.stderr.write(r.stderr)
                raise SystemExit(1)
            return o, r.stderr

        with ThreadPoolExecutor(max_workers=8) as ex:
            res = list(ex.map(one, B.SOURCES))
        for (o, log), src in zip(res, B.SOURCES):
            keep = False
            for line in log.splitlines():
                if "Compiling entry function" in line:
                    keep = ("k_count" in line and "prep" not in line) or "k_od_count" in line
                elif keep and "registers" in line:
                    print(name, src, line.strip())
        lib = os.path.join(out, f"libfofb200_{name}.so")
        subprocess.run([nvcc, "-shared", "-o", lib] + [o for o, _ in res] + ["-lcudart"], check=True)
        print("built", lib)


if __name__ == "__main__":
    main()
