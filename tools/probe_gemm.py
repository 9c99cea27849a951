"""GPU bring-up probe for the tcgen05 GEMM: every case runs in its own subprocess (a device
trap poisons the CUDA context), results go to gpurun_out/probe_gemm.jsonl.

    python tools/probe_gemm.py            # all cases
    python tools/probe_gemm.py --case tn_bf16
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
OUT = ROOT / "gpurun_out"


def rel_err(got, ref):
    import torch
    return float((got.double() - ref.double()).abs().max() / ref.double().abs().max().clamp_min(1e-30))


def run_case(name: str) -> dict:
    import torch
    from sdumc_b200 import ops
    torch.manual_seed(0)
    dev = "cuda"
    res = {"case": name}

    def mk(shape, dtype):
        return (torch.randn(*shape, device=dev) * 0.5).to(dtype).contiguous()

    def ref_mm(A, B, a_mn, b_mn):
        a = A.float().t() if a_mn else A.float()       # [M,K]
        b = B.float() if b_mn else B.float().t()       # [K,N]
        return a.double() @ b.double()

    kind, *rest = name.split(":")
    if kind == "mm":
        # mm:<dtype>:<a_mn><b_mn>:M:N:K:block_n:ksplit[:lbo:sbo]
        dt = torch.bfloat16 if rest[0] == "bf16" else torch.float32
        a_mn, b_mn = int(rest[1][0]), int(rest[1][1])
        M, N, K, bn, ks = map(int, rest[2:7])
        lbo = int(rest[7]) if len(rest) > 7 else 0
        sbo = int(rest[8]) if len(rest) > 8 else 0
        A = mk((K, M) if a_mn else (M, K), dt)
        B = mk((K, N) if b_mn else (N, K), dt)
        bias = torch.randn(N, device=dev) if ks == 1 else None
        out = torch.zeros(M, N, device=dev)
        ops.gemm(A, B, M=M, N=N, K=K, a_mn=a_mn, b_mn=b_mn, k_splits=ks, block_n=bn, bias=bias, out_f32=out,
                 f32_mode=ops.OUT_ATOMIC if ks > 1 else ops.OUT_STORE)
        torch.cuda.synchronize()
        ref = ref_mm(A, B, a_mn, b_mn)
        if dt == torch.float32:  # tf32 truncation of the operands
            pass
        if bias is not None:
            ref = ref + bias.double()
        res["rel_err"] = rel_err(out, ref)
        res["ok"] = res["rel_err"] < (2e-3 if dt == torch.float32 else 1e-3)
    elif kind == "inproj":
        M, N, K = 1000, 256, 1024
        A, B = mk((M, K), torch.bfloat16), mk((N, K), torch.bfloat16)
        bias = torch.randn(N, device=dev)
        H = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        tg = [torch.zeros(M, N, device=dev, dtype=torch.bfloat16) for _ in range(4)]
        sites = [11, 12, 13, 14]
        ops.gemm(A, B, M=M, N=N, K=K, bias=bias, out_bf16=H, epi_kind=ops.EPI_INPROJ, targets=tg, target_sites=sites,
                 seed=0x1234567890, step=7)
        torch.cuda.synchronize()
        ref = ref_mm(A, B, 0, 0) + bias.double()
        res["rel_err"] = rel_err(H, ref)
        errs = []
        keep = []
        for t, s in zip(tg, sites):
            mask = ops.frame_mask(0x1234567890, 7, s, M, N)
            errs.append(rel_err(t, (H.float() * mask).bfloat16().float()))
            keep.append(float((mask > 0).float().mean()))
        res["mask_err"] = errs
        res["keep_frac"] = keep
        res["ok"] = res["rel_err"] < 1e-2 and max(errs) < 1e-6 and all(0.49 < k < 0.51 for k in keep)
    elif kind == "keyproj":
        nq = int(rest[0])
        Bn, L, G = 37, 50, 256
        M = Bn * L
        A, W = mk((M, G), torch.bfloat16), (torch.randn(G, G, device=dev) * 0.06).bfloat16()
        bias = torch.randn(G, device=dev) * 0.1
        q = torch.randn(Bn if nq > 1 else 1, nq, G, device=dev)
        S = torch.zeros(M, nq, device=dev)
        Kout = torch.zeros(M, G, device=dev, dtype=torch.bfloat16)
        ops.gemm(A, W, M=M, N=G, K=G, bias=bias, act=ops.ACT_TANH, epi_kind=ops.EPI_KEYPROJ, out_bf16=Kout, qv=q,
                 q_stride=(nq * G if nq > 1 else 0), nq=nq, L=L, scores=S)
        torch.cuda.synchronize()
        Kref = torch.tanh(ref_mm(A, W, 0, 0) + bias.double())
        res["k_err"] = rel_err(Kout, Kref)
        qq = q.double() if nq > 1 else q.double().expand(Bn, nq, G)
        Sref = torch.einsum("blg,bqg->blq", Kout.double().view(Bn, L, G), qq).reshape(M, nq)
        res["s_err"] = rel_err(S, Sref)
        res["ok"] = res["k_err"] < 1e-2 and res["s_err"] < 1e-4
    elif kind == "epi":
        # generic epilogue options: relu + element dropout, gate, fmask, += modes, bf16 RMW
        M, N, K = 515, 192, 256
        A, B = mk((M, K), torch.float32), mk((N, K), torch.float32)
        bias = torch.randn(N, device=dev)
        out = torch.zeros(M, N, device=dev)
        ops.gemm(A, B, M=M, N=N, K=K, bias=bias, act=ops.ACT_RELU, drop_p=0.3, drop_site=5, out_f32=out, seed=99,
                 step=3)
        mask = ops.elem_mask(99, 3, 5, M * N, 0.3).view(M, N)
        ref = torch.relu(ref_mm(A, B, 0, 0) + bias.double()) * mask.double()
        res["drop_err"] = rel_err(out, ref)
        res["drop_keep"] = float((mask > 0).float().mean())
        gate = torch.randn(M, N, device=dev)
        prev = torch.randn(M, N, device=dev)
        out2 = prev.clone()
        ops.gemm(A, B, M=M, N=N, K=K, gate=gate, gate_scale=1.0 / 0.7, out_f32=out2, f32_mode=ops.OUT_ADD)
        ref2 = prev.double() + ref_mm(A, B, 0, 0) * (gate > 0).double() / 0.7
        res["gate_add_err"] = rel_err(out2, ref2)
        Nn = 256
        B3 = mk((Nn, K), torch.float32)
        prevb = torch.randn(M, Nn, device=dev).bfloat16()
        out3 = prevb.clone()
        ops.gemm(A, B3, M=M, N=Nn, K=K, fmask_site=21, out_bf16=out3, bf16_mode=ops.OUT_ADD, seed=99, step=3)
        fm = ops.frame_mask(99, 3, 21, M, Nn)
        ref3 = prevb.double() + ref_mm(A, B3, 0, 0) * fm.double()
        res["fmask_rmw_err"] = rel_err(out3, ref3)
        torch.cuda.synchronize()
        res["ok"] = res["drop_err"] < 2e-3 and res["gate_add_err"] < 2e-3 and res["fmask_rmw_err"] < 1e-2 and \
            0.68 < res["drop_keep"] < 0.72
    elif kind == "time":
        # time:<M>:<N>:<K>:<a_mn><b_mn>:<ksplit>
        M, N, K = map(int, rest[0:3])
        a_mn, b_mn = int(rest[3][0]), int(rest[3][1])
        ks = int(rest[4])
        noout = len(rest) > 5 and rest[5] == "noout"
        bn = int(rest[6]) if len(rest) > 6 else 0
        dbg = int(rest[7]) if len(rest) > 7 else 0
        dt = torch.float32 if (len(rest) > 8 and rest[8] == "f32") else torch.bfloat16   # f32: tf32 MMAs, fp32 + bf16 outputs
        A = mk((K, M) if a_mn else (M, K), dt)
        B = mk((K, N) if b_mn else (N, K), dt)
        outb = torch.zeros(M, N, device=dev, dtype=torch.bfloat16) if ks == 1 else None
        outf = torch.zeros(M, N, device=dev) if (ks > 1 or dt == torch.float32) else None
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

        def go():
            ops.gemm(A, B, M=M, N=N, K=K, a_mn=a_mn, b_mn=b_mn, k_splits=ks, out_bf16=None if noout else outb,
                     out_f32=None if noout else outf, block_n=bn, max_ctas=-1 if dbg == 1 else 0,   # -1: no CTA pairs
                     f32_mode=ops.OUT_ATOMIC if ks > 1 else ops.OUT_STORE)
        for _ in range(3):
            go()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            go()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        res["ms"] = ms
        res["tflops"] = 2.0 * M * N * K / ms / 1e9
        res["gbs"] = (A.numel() * 2 + B.numel() * 2 + M * N * (2 if ks == 1 else 4)) / ms / 1e6
        res["ok"] = True
    elif kind == "memref":
        # memref:<MB>  torch reference rates on the same box: read-only (sum), copy, write-only (fill), L2 flushed
        mb = int(rest[0])
        x = torch.randn(mb << 19, device=dev, dtype=torch.bfloat16)
        y = torch.empty_like(x)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

        def t(fn):
            ts = []
            for _ in range(3):
                fn()
            for _ in range(10):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            return sorted(ts)[len(ts) // 2]
        xi = x.view(torch.int32)
        res["sum_ms"] = t(lambda: xi.sum())
        res["sum_gbs"] = (mb << 20) / res["sum_ms"] / 1e6
        res["copy_ms"] = t(lambda: y.copy_(x))
        res["copy_gbs"] = 2 * (mb << 20) / res["copy_ms"] / 1e6
        res["fill_ms"] = t(lambda: y.zero_())
        res["fill_gbs"] = (mb << 20) / res["fill_ms"] / 1e6
        res["ok"] = True
    elif kind == "nullchain":
        # 20 dependent trivial kernels (cast of 1 KB) replayed as a CUDA graph: the floor of a dependent launch
        a = torch.zeros(256, device=dev)
        b = torch.zeros(256, device=dev, dtype=torch.bfloat16)

        def body():
            for _ in range(20):
                ops.cast_bf16(a, b)
        body()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            body()
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res["us_per_kernel"] = sorted(ts)[len(ts) // 2] * 1e3 / 20
        res["ok"] = True
    elif kind == "chain":
        # chain:<M>:<N>:<K>:<bn>:<cold>  20 dependent tf32 linears (bias + ReLU + dropout) replayed as a CUDA graph
        M, N, K, bn, cold = map(int, rest[0:5])
        lvl = int(rest[5]) if len(rest) > 5 else 3     # epilogue level: 0 plain store, 1 +bias/ReLU, 2 +dropout, 3 +bf16 copy
        Ws = [mk((N, K), torch.float32) * 0.1 for _ in range(20)]
        bs_ = [torch.randn(N, device=dev) for _ in range(20)]
        xs = [mk((M, K), torch.float32), mk((M, K), torch.float32)]
        xb = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

        def body():
            for i in range(20):
                ops.gemm(xs[i & 1], Ws[i], M=M, N=N, K=K, bias=bs_[i] if lvl >= 1 else None,
                         act=ops.ACT_RELU if lvl >= 1 else ops.ACT_NONE, drop_p=0.3 if lvl >= 2 else 0.0, drop_site=3 + i,
                         out_f32=xs[(i + 1) & 1], out_bf16=xb if lvl >= 3 else None, seed=5, step=1, block_n=bn)
        body()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            body()
        ts = []
        for _ in range(10):
            if cold:
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res["us_per_layer"] = sorted(ts)[len(ts) // 2] * 1e3 / 20
        res["ok"] = True
    elif kind == "timeepi":
        # timeepi:inproj:<ntgt>:<K>  |  timeepi:keyproj:<nq>:<store_k>
        sub = rest[0]
        if sub == "rmw":
            # dH += (dZ W_in) * mask : M=196608, N=K=256, B MN-major, bf16 read-modify-write epilogue
            masked = int(rest[1])
            M, N, K = 196608, 256, 256
            A, B = mk((M, K), torch.bfloat16), mk((K, N), torch.bfloat16)
            H = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)

            def go():
                ops.gemm(A, B, M=M, N=N, K=K, b_mn=True, fmask_site=7 if masked else 0, out_bf16=H,
                         bf16_mode=ops.OUT_ADD, seed=5, step=1)
            nbytes = A.numel() * 2 + 2 * M * N * 2
        elif sub == "inproj":
            ntgt, K = int(rest[1]), int(rest[2])
            M, N = 196608, 256
            A, B = mk((M, K), torch.bfloat16), mk((N, K), torch.bfloat16)
            bias = torch.randn(N, device=dev)
            H = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
            tg = [torch.zeros(M, N, device=dev, dtype=torch.bfloat16) for _ in range(ntgt)]

            def go():
                ops.gemm(A, B, M=M, N=N, K=K, bias=bias, out_bf16=H if ntgt == 0 else None, epi_kind=ops.EPI_INPROJ,
                         targets=tg, target_sites=list(range(1, ntgt + 1)), seed=5, step=1)
            nbytes = A.numel() * 2 + M * N * 2 * max(1, ntgt)
        else:
            nq, store_k = int(rest[1]), int(rest[2])
            L, G = 384, 256
            M = 512 * L
            K = G
            N = G
            A, B = mk((M, G), torch.bfloat16), (torch.randn(G, G, device=dev) * 0.06).bfloat16()
            bias = torch.randn(G, device=dev) * 0.1
            q = torch.randn(512 if nq > 1 else 1, nq, G, device=dev)
            S = torch.zeros(M, nq, device=dev)
            Kout = torch.zeros(M, G, device=dev, dtype=torch.bfloat16) if store_k else None

            def go():
                ops.gemm(A, B, M=M, N=G, K=G, bias=bias, act=ops.ACT_TANH, epi_kind=ops.EPI_KEYPROJ, out_bf16=Kout,
                         qv=q, q_stride=(nq * G if nq > 1 else 0), nq=nq, L=L, scores=S)
            nbytes = A.numel() * 2 + (M * G * 2 if store_k else 0) + M * nq * 4
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        for _ in range(3):
            go()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            go()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        res["ms"] = ms
        res["tflops"] = 2.0 * M * N * K / ms / 1e9
        res["gbs"] = nbytes / ms / 1e6
        res["ok"] = True
    else:
        raise ValueError(name)
    return res


CASES = [
    "mm:bf16:00:300:256:1024:256:1",
    "mm:bf16:00:300:256:1024:128:1",
    "mm:bf16:00:300:200:1024:64:1",
    "mm:bf16:00:1000:256:4096:0:1",
    "mm:f32:00:300:256:512:256:1",
    "mm:f32:00:300:128:896:64:1",
    # MN-major B (dX = dY W)
    "mm:bf16:01:300:256:256:256:1",
    "mm:bf16:01:300:256:256:64:1",
    # MN-major A
    "mm:bf16:10:256:256:1000:128:1",
    # both MN-major (dW = dY^T X), with split-K
    "mm:bf16:11:256:1024:5000:256:1",
    "mm:bf16:11:256:1024:5000:256:8",
    "inproj",
    "keyproj:1",
    "keyproj:7",
    "epi",
    "time:196608:256:1024:00:1",
    "time:32768:256:4096:00:1",
    "time:256:1024:196608:11:37",
    "time:196608:256:256:00:1",
    "timeepi:rmw:1",
    "timeepi:rmw:0",
    "timeepi:inproj:0:1024",
    "timeepi:inproj:4:1024",
    "timeepi:inproj:0:4096",
    "timeepi:keyproj:1:0",
    "timeepi:keyproj:1:1",
    "timeepi:keyproj:7:0",
    "timeepi:keyproj:7:1",
]


def main():
    if "--case" in sys.argv:
        name = sys.argv[sys.argv.index("--case") + 1]
        try:
            r = run_case(name)
        except Exception as e:  # noqa: BLE001
            r = {"case": name, "ok": False, "error": f"{type(e).__name__}: {e}"[:500]}
        print("RESULT " + json.dumps(r), flush=True)
        return
    OUT.mkdir(exist_ok=True)
    log = open(OUT / "probe_gemm.jsonl", "w")
    n_bad = 0
    cases = sys.argv[sys.argv.index("--cases") + 1].split(",") if "--cases" in sys.argv else CASES
    for c in cases:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, __file__, "--case", c], capture_output=True, text=True, timeout=180)
            line = next((l for l in p.stdout.splitlines() if l.startswith("RESULT ")), None)
            r = json.loads(line[7:]) if line else {"case": c, "ok": False, "error": "no result",
                                                   "stderr": p.stderr[-800:], "rc": p.returncode}
        except subprocess.TimeoutExpired:
            r = {"case": c, "ok": False, "error": "timeout"}
        r["wall_s"] = round(time.time() - t0, 1)
        n_bad += 0 if r.get("ok") else 1
        s = json.dumps(r)
        print(s, flush=True)
        log.write(s + "\n")
        log.flush()
    print(f"probe_gemm: {len(cases) - n_bad}/{len(cases)} ok")


if __name__ == "__main__":
    main()
