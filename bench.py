#!/usr/bin/env python
"""bench.py — the headline metric of BASELINE.json on its own configuration.

metric   : batched LZ decode, decompressed GB/s (LZ10), and fraction of the measured HBM roofline
workload : config C2 — 65 536 synthetic 64 KiB GBA/DS-style asset streams (classes T/M/X, 40/30/30 %), LZ10,
           encoded at quality 8 / VRAM mode on (the reference defaults) by this engine's GPU encoder, which is
           byte-identical to the reference encoder (tests/test_encode_gpu.py)
step     : one decode pass over the whole batch
value    : decompressed bytes / device time of the decode kernel with inputs already resident in HBM
e2e      : the same batch through aurora_decode_batch with pinned HOST buffers (H2D + kernel + D2H timed)
N > 1    : one process per GPU (torchrun); every rank decodes its own 65 536-stream shard (weak scaling, no
           collective on the data path); the time is the max over ranks, value the sum of bytes / that time

`--impl reference` times the reference's CPU implementation of the same path — the C++ restatement in oracle/
(the reference is managed C# and there is no .NET runtime on the box; DESIGN.md "Oracle") — on all host threads.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "batched LZ10 decode throughput (decompressed GB/s out)"
UNIT = "GB/s"
STREAM_BYTES = 65536
FULL_STREAMS = 65536
QUALITY = 8


def workload_config(streams):
    return {
        "workload": "C2: LZ10 batched decode, 65536 synthetic 64 KiB GBA/DS-style asset streams per GPU",
        "format": "LZ10",
        "streams_per_gpu": streams,
        "stream_bytes": STREAM_BYTES,
        "classes": "T 40% / M 30% / X 30% (SURVEY.md 8d)",
        "encoder": "quality 8, GbaVramCompatibilityMode on (reference defaults)",
        "l2_policy": "inputs larger than L2 (compressed + decompressed working set >> 126 MB per step)",
    }


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                clk, mxc = float(f[1]), float(f[2])
            except ValueError:
                continue
            mx = mxc
            if t0 - 0.05 <= ts <= t1 + 0.15:
                sm.append(clk)
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        if not sm:
            sm = [float(l.split(",")[1]) for _, l in self.lines[-3:] if len(l.split(",")) > 2] or [0.0]
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def ncu_traffic(streams):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the decode kernel, from the committed ncu capture."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        e = d.get("lz10_c2")
        if e and e.get("streams") == streams:
            return float(e["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np
    import torch
    from auroralib.compression_b200 import _abi as A, corpus
    from auroralib.compression_b200.batch import layout
    from oracle import oracle as O
    O.build()
    threads = O.hardware_threads()
    sample = min(args.streams, 8192)   # bounded sample of the C2 workload per step (512 MiB decoded)
    raw, _ = corpus.generate_mix(sample, STREAM_BYTES, device="cpu")
    raw_h = raw.numpy().reshape(-1)
    roff = np.arange(sample, dtype=np.uint64) * np.uint64(STREAM_BYTES)
    rlen = np.full(sample, STREAM_BYTES, dtype=np.uint64)
    caps, coff, ctotal = layout([STREAM_BYTES + STREAM_BYTES // 8 + 64] * sample)
    comp = np.zeros(ctotal + 16, dtype=np.uint8)
    clen, st = O.encode_packed(A.FMT_LZ10, raw_h, roff, rlen, comp, coff, caps, A.make_opts(quality=QUALITY), threads)
    assert (st == 0).all()
    dst = np.zeros(sample * STREAM_BYTES + 16, dtype=np.uint8)
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        out_len, consumed, status = O.decode_packed(A.FMT_LZ10, comp, coff, clen, dst, roff, rlen, None, threads)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    assert (status == 0).all() and np.array_equal(dst[:sample * STREAM_BYTES], raw_h)
    total = sum(times)
    value = sample * STREAM_BYTES * len(times) / total / 1e9
    sample_desc = f"{sample} of the {args.streams} C2 streams per step ({sample * STREAM_BYTES >> 20} MiB decoded), std::thread over {threads} host threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(1000 * total / len(times), 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": workload_config(args.streams),
        "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": threads, "kind": "port", "sample": sample_desc,
                         "note": "C++ restatement of the reference's managed decoder (no .NET runtime on the box)"},
        "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from auroralib.compression_b200 import BatchCodec, _abi as A, corpus, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    codec = BatchCodec(device_mask=1 << local_rank)
    n, size = args.streams, STREAM_BYTES
    fmt = A.FMT_LZ10
    opts = A.make_opts(quality=QUALITY)

    # ---- synthetic corpus on the device, encoded by the engine's own GPU encoder (byte-identical to the reference's)
    raw, classes = corpus.generate_mix(n, size, seed=0xA0120000 + 7919 * rank, device=dev)
    bound = codec.encode_bound(fmt, size)
    bound16 = (bound + 15) & ~15
    i64 = dict(dtype=torch.int64, device=dev)
    r_off = torch.arange(n, **i64) * size
    r_len = torch.full((n,), size, **i64)
    c_buf = torch.empty(n * bound16 + 16, dtype=torch.uint8, device=dev)
    c_off = torch.arange(n, **i64) * bound16
    c_cap = torch.full((n,), bound, **i64)
    c_len = torch.zeros(n, **i64)
    e_st = torch.zeros(n, dtype=torch.int32, device=dev)
    ts = torch.cuda.Stream(device=dev)
    t_enc0 = time.perf_counter()
    codec.encode_device(fmt, raw.view(-1), r_off, r_len, c_buf, c_off, c_cap, c_len, e_st, opts, stream=ts.cuda_stream)
    ts.synchronize()
    t_enc = time.perf_counter() - t_enc0
    assert int(e_st.abs().sum()) == 0, "GPU encode failed"
    # pack the compressed streams tightly (16-byte aligned), the layout a real batch would have
    pad = (c_len + 15) & ~15
    p_off = torch.cumsum(pad, 0) - pad
    comp_total = int(pad.sum())
    packed = torch.zeros(comp_total + 16, dtype=torch.uint8, device=dev)
    idx = torch.arange(bound16, device=dev).unsqueeze(0)
    for s in range(0, n, 4096):
        e = min(n, s + 4096)
        src = c_buf[s * bound16:e * bound16].view(e - s, bound16)
        mask = idx < c_len[s:e].unsqueeze(1)
        pos = (p_off[s:e].unsqueeze(1) + idx)[mask]
        packed[pos] = src[mask]
    del c_buf, idx, mask, pos, src
    torch.cuda.empty_cache()
    comp_bytes = int(c_len.sum())
    out_bytes = n * size

    d_dst = torch.zeros(out_bytes + 16, dtype=torch.uint8, device=dev)
    d_olen = torch.zeros(n, **i64)
    d_cons = torch.zeros(n, **i64)
    d_st = torch.zeros(n, dtype=torch.int32, device=dev)

    def decode_once():
        codec.decode_device(fmt, packed, p_off, c_len, d_dst, r_off, r_len, d_olen, d_cons, d_st, stream=ts.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: W warm-ups, K timed steps, CUDA events on the launching stream
    for _ in range(max(args.warmup, 3)):
        decode_once()
    ts.synchronize()
    assert int(d_st.abs().sum()) == 0 and torch.equal(d_dst[:out_bytes].view(n, size), raw), "decode mismatch"
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    barrier()
    launches0 = codec.kernel_launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_wall0 = time.time()
    with torch.cuda.stream(ts):
        for a, b in ev:
            a.record(ts)
            decode_once()
            b.record(ts)
    barrier()
    t_wall1 = time.time()
    launches = codec.kernel_launches - launches0
    clocks = sampler.stop(t_wall0, t_wall1)
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = ev[0][0].elapsed_time(ev[-1][1])
    kernel_ms = sum(step_ms) / len(step_ms)
    tmax = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms_max = float(tmax.item())
    ms_per_step = total_ms_max / args.steps
    value = world * out_bytes / (ms_per_step * 1e-3) / 1e9

    # ---- per-class device timing (outside the headline region; explains the mix)
    per_class = {}
    if rank == 0:
        cls_idx = {c: torch.tensor([i for i, k in enumerate(classes) if k == c], device=dev) for c in sorted(set(classes))}
        for c, ix in cls_idx.items():
            po, cl, ro, rl = p_off[ix].contiguous(), c_len[ix].contiguous(), r_off[ix].contiguous(), r_len[ix].contiguous()
            ol, co = torch.zeros(len(ix), **i64), torch.zeros(len(ix), **i64)
            st_ = torch.zeros(len(ix), dtype=torch.int32, device=dev)
            best = 1e30
            for _ in range(3):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                with torch.cuda.stream(ts):
                    a.record(ts)
                    codec.decode_device(fmt, packed, po, cl, d_dst, ro, rl, ol, co, st_, stream=ts.cuda_stream)
                    b.record(ts)
                ts.synchronize()
                best = min(best, a.elapsed_time(b))
            cb, ob = int(cl.sum()), len(ix) * size
            per_class[c] = {"streams": len(ix), "ratio": round(cb / ob, 4), "out_gbs": round(ob / best / 1e6, 1),
                            "hbm_gbs": round((cb + ob) / best / 1e6, 1)}

    # ---- end to end through the C ABI with pinned host buffers (H2D + kernel + D2H inside the timed region)
    L = _lib.load()
    h_src_p = L.aurora_pinned_alloc(comp_total + 16)
    h_dst_p = L.aurora_pinned_alloc(out_bytes + 16)
    assert h_src_p and h_dst_p, "pinned allocation failed"
    h_src = np.ctypeslib.as_array(C.cast(h_src_p, C.POINTER(C.c_uint8)), shape=(comp_total + 16,))
    h_dst = np.ctypeslib.as_array(C.cast(h_dst_p, C.POINTER(C.c_uint8)), shape=(out_bytes + 16,))
    torch.from_numpy(h_src).copy_(packed)
    h_off = p_off.cpu().numpy().astype(np.uint64)
    h_len = c_len.cpu().numpy().astype(np.uint64)
    h_doff = r_off.cpu().numpy().astype(np.uint64)
    h_cap = r_len.cpu().numpy().astype(np.uint64)
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(2):
        out_len, consumed, status = codec.decode_packed(fmt, h_src, h_off, h_len, h_dst, h_doff, h_cap, opts)
    assert (status == 0).all()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        out_len, consumed, status = codec.decode_packed(fmt, h_src, h_off, h_len, h_dst, h_doff, h_cap, opts)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    e2e_ok = bool((status == 0).all()) and bool(np.array_equal(h_dst[:out_bytes], raw.view(-1).cpu().numpy()))
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * out_bytes / float(te.item()) / 1e9
    h2d = comp_total + 4 * 8 * n
    d2h = out_bytes + n * (8 + 8 + 4)

    # ---- CPU baseline (rank 0, N == 1 only): the oracle on a bounded sample of the same streams
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        O.build()
        threads = O.hardware_threads()
        sample = min(n, 16384)
        cdst = np.zeros(sample * size + 16, dtype=np.uint8)
        best = 1e30
        for _ in range(3):
            t0 = time.perf_counter()
            ol, co, st_ = O.decode_packed(fmt, h_src, h_off[:sample], h_len[:sample], cdst, h_doff[:sample], h_cap[:sample], None, threads)
            best = min(best, time.perf_counter() - t0)
        assert (st_ == 0).all() and np.array_equal(cdst[:sample * size], h_dst[:sample * size]), "oracle and GPU disagree"
        cpu_baseline = {"value": round(sample * size / best / 1e9, 3), "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"first {sample} of the {n} streams ({sample * size >> 20} MiB decoded), best of 3, std::thread over {threads} host threads",
                        "note": "C++ restatement of the reference's managed decoder (no .NET runtime on the box)"}
    L.aurora_pinned_free(h_src_p)
    L.aurora_pinned_free(h_dst_p)

    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = (comp_bytes + out_bytes) / (kernel_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": workload_config(n),
            "roofline": {"bound": "hbm", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                         "traffic": ncu_traffic(n), "peak_source": peak_src, "kernel": "decode_flaglz_kernel<LZ10>",
                         "algorithmic_bytes_per_launch": comp_bytes + out_bytes, "kernel_ms": round(kernel_ms, 4)},
            "cpu_baseline": cpu_baseline,
            "e2e": {"value": round(e2e_value, 3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "verified": e2e_ok, "steps": e2e_steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "per_class": per_class,
            "compression_ratio": round(comp_bytes / out_bytes, 4),
            "encode_s": round(t_enc, 3),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    codec.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=FULL_STREAMS, help="streams per GPU (default: the full C2 configuration)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
