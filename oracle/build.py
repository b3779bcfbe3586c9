"""Builds the C/OpenMP oracle restatement (oracle/esac_oracle_c.c) into oracle/_build/ (git-ignored).

The reference extension itself is NOT buildable in this image (esac.cpp:32 needs the OpenCV C++ headers and
libopencv_core / libopencv_calib3d; only the cv2 Python wheel exists), so there is no oracle/_ref/."""
from __future__ import annotations

import ctypes as C
import hashlib
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
SRC = HERE / "esac_oracle_c.c"
OUT = HERE / "_build" / "libesac_oracle.so"


def build_oracle(force: bool = False) -> Path:
    OUT.parent.mkdir(exist_ok=True)
    stamp = OUT.parent / "libesac_oracle.so.srchash"
    want = hashlib.sha256(SRC.read_bytes()).hexdigest()
    if force or not OUT.exists() or not stamp.exists() or stamp.read_text().strip() != want:
        cmd = ["/usr/bin/gcc", "-O3", "-fopenmp", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(OUT), str(SRC), "-lm"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"gcc failed:\n{r.stdout}\n{r.stderr}")
        stamp.write_text(want)
    return OUT


def load_oracle() -> C.CDLL:
    lib = C.CDLL(str(build_oracle()))
    vp, i32, f32 = C.c_void_p, C.c_int, C.c_float
    lib.esac_oracle_score.argtypes = [vp, i32, i32, i32, vp, i32, vp, i32, i32, f32, f32, f32, f32, f32, f32, f32, i32, vp, i32]
    lib.esac_oracle_score.restype = i32
    lib.esac_oracle_max_threads.restype = i32
    return lib


def c_score(coords, assign, poses6, shiftX, shiftY, f, ppx, ppy, tau, alpha, beta, max_reproj, sub, nthreads=0):
    """Scores [M] from the C restatement; returns (scores, threads used)."""
    import numpy as np
    lib = load_oracle()
    coords = np.ascontiguousarray(coords, np.float32)
    assign = np.ascontiguousarray(assign, np.int64)
    poses6 = np.ascontiguousarray(poses6, np.float64)
    E, _, H, W = coords.shape
    out = np.zeros(len(assign))
    used = lib.esac_oracle_score(coords.ctypes.data, E, H, W, assign.ctypes.data, len(assign), poses6.ctypes.data,
                                 int(shiftX), int(shiftY), float(f), float(ppx), float(ppy), float(tau), float(alpha),
                                 float(beta), float(max_reproj), int(sub), out.ctypes.data, int(nthreads))
    return out, used


if __name__ == "__main__":
    print(build_oracle(force=True))
