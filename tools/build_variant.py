"""Build an experimental variant of libmicrodit_b200.so with extra -D flags on the GEMM (A/B kernel experiments):

    python tools/build_variant.py epi4 -DMD_EPI_WARPS=4      -> variants/libmicrodit_b200_epi4.so
    MD_LIB_PATH=variants/libmicrodit_b200_epi4.so python tools/gemm_micro.py
"""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from micro_diffusion_b200 import build as B  # noqa: E402

name, flags = sys.argv[1], sys.argv[2:]
B.build()
out = ROOT / "variants"
out.mkdir(exist_ok=True)
obj = out / f"gemm_{name}.o"
subprocess.run([B.NVCC, *B.FLAGS, *flags, "-c", str(B.CSRC / "gemm_tcgen05.cu"), "-o", str(obj)], check=True)
others = [str(o) for o in (B.CSRC / "build").glob("*.o") if o.stem != "gemm_tcgen05"]
lib = out / f"libmicrodit_b200_{name}.so"
subprocess.run([B.NVCC, "-shared", "-cudart", "static", "-o", str(lib), str(obj), *others], check=True)
print(lib)
