#!/usr/bin/env python
"""Measure the SM-side ceilings that bound the streaming kernels (fp64 DFMA rate and latency, L2 -> SM read
bandwidth with 256-bit loads, shared-memory read bandwidth) and write them to profiles/sm_peaks.json.

    python scripts/measure_sm_peaks.py            # on a B200 box; builds mcmc_spec_b200/sm_peaks if missing

MEASURED_PEAKS.json (driver-written) only has the HBM copy and the bf16 GEMM peaks; neither binds this path.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, 'mcmc_spec_b200', 'sm_peaks')
SRC = os.path.join(ROOT, 'mcmc_spec_b200', 'csrc', 'sm_peaks.cu')


def build():
    if os.path.exists(BIN) and os.path.getmtime(BIN) >= os.path.getmtime(SRC):
        return
    nvcc = os.environ.get('NVCC') or '/usr/local/cuda/bin/nvcc'
    subprocess.run([nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-o', BIN, SRC], check=True)


def main():
    build()
    out = subprocess.run([BIN], check=True, capture_output=True, text=True).stdout
    data = json.loads(out)
    dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'profiles', 'sm_peaks.json')
    with open(dst, 'w') as f:
        json.dump(data, f, indent=1)
    print(json.dumps(data))


if __name__ == '__main__':
    main()
