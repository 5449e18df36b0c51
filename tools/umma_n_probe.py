"""tcgen05 (cta_group::2, M=256) throughput as a function of the instruction's N: N = tiles x W
with the 2-CTA kernel's even split, so every UMMA is exactly W wide; several n-tiles per row block
keep the A traffic off the critical path.  CUDA events, isolated kernel."""
import json, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from reprover_b200 import _native

lib = _native.load()
dev = torch.device("cuda:0")
out = {}
M, K = 74 * 256 * 4, 4096
A = torch.randn(M, K, device=dev).to(torch.bfloat16)
st = torch.cuda.current_stream().cuda_stream
for N, W in [(480, 160), (576, 192), (896, 224), (768, 256), (1024, 256), (1568, 224), (1792, 256)]:  # N = tiles x W
    B = torch.randn(N, K, device=dev).to(torch.bfloat16)
    C = torch.empty(M, N, device=dev)
    for _ in range(3):
        _native.check(lib.rpx_gemm2_bf16_f32(A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, st))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        lib.rpx_gemm2_bf16_f32(A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    out[N] = {"tile_width": W, "ms": ms, "tflops": 2.0 * M * N * K / ms / 1e9}
    print(N, json.dumps(out[N]), flush=True)
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/umma_n_probe.json").write_text(json.dumps(out, indent=1))
