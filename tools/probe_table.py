"""Compact table of a tools/gemm_probe.py output file. usage: python tools/probe_table.py file [mode]"""
import json
import sys

mode = sys.argv[2] if len(sys.argv) > 2 else "1"
for l in open(sys.argv[1]):
    try:
        d = json.loads(l)
    except Exception:  # noqa: BLE001
        continue
    k = f"emx_mode{mode}"
    print(f"{d['batch']:3d} {d['gemm']:14s} M={d['M']:5d} N={d['N']:5d} K={d['K']:5d} {d['epilogue']:12s} cublas {d['cublas_plain_ms'] * 1e3:7.1f}us "
          f"{d['cublas_tflops']:7.1f}  emx {d[k + '_ms'] * 1e3:7.1f}us {d[k + '_tflops']:7.1f}  ratio {d[k + '_tflops'] / d['cublas_tflops']:.2f}")
