"""Large-graph walk: tcgen05 GEMM path vs exact-fp32 SIMT path, and the raw batched GEMM against torch.bmm (cuBLAS sgemm).
One JSON object per line; CUDA events, 5 warm-ups, median of 20."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sapienza_video_contrastive_b200 import ops  # noqa: E402
from tools.sweep import timeit  # noqa: E402

dev = torch.device("cuda", 0)


def main():
    torch.backends.cuda.matmul.allow_tf32 = False
    for (Z, M, N, K) in [(24, 128, 128, 128), (24, 256, 256, 256), (12, 512, 512, 512), (6, 1024, 1024, 1024), (24, 1024, 128, 1024)]:
        A = torch.randn(Z, M, K, device=dev)
        B = torch.randn(Z, K, N, device=dev)
        out = torch.empty(Z, M, N, device=dev)
        t_tc = timeit(lambda: ops.bmm_tc(A, B, out=out))
        t_tf32 = timeit(lambda: ops.bmm_tf32(A, B, out=out))
        t_cublas = timeit(lambda: torch.bmm(A, B, out=out))
        fl = 2.0 * Z * M * N * K
        print(json.dumps({"kind": "bmm", "Z": Z, "M": M, "N": N, "K": K, "tc_ms": t_tc, "tf32_fused_ms": t_tf32, "cublas_sgemm_ms": t_cublas,
                          "tc_tflops": fl / t_tc / 1e9, "tf32_fused_tflops": fl / t_tf32 / 1e9, "cublas_tflops": fl / t_cublas / 1e9}), flush=True)
    ones = torch.ones(1, device=dev)
    for (B, N, T) in [(8, 128, 4), (8, 196, 8), (8, 256, 4), (4, 512, 4), (2, 1024, 4), (4, 256, 8), (1, 1024, 8)]:
        f = torch.randn(B, N, T, 128, device=dev, requires_grad=True)
        row = {"kind": "walk_fwd_bwd", "B": B, "N": N, "T": T}
        for name, kw in (("simt_ms", dict(force_simt=True)), ("f16split_ms", dict(no_tf32=True)), ("tc_ms", dict())):
            def step():
                f.grad = None
                q, loss, xent, acc = ops.walk(f, 0.07, 0.1, rng="philox", **kw)
                loss.backward(ones)
            row[name] = timeit(step)
        flops = 3 * (2 * (T - 1) * N * N * 128 + 2 * N ** 3 * 3 * (T - 2)) * B
        row["tc_algorithmic_tflops"] = flops / row["tc_ms"] / 1e9
        row["speedup"] = row["simt_ms"] / row["tc_ms"]
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
