#!/usr/bin/env python
"""bench.py -- 1080p H.264 macroblock reconstruction throughput (BASELINE.json: "1080p recon
frames/s per GPU and 8-GPU GOP-sharded; HBM GB/s vs B200 peak").

Workload (config.workload): 1920x1080 (120x68 macroblocks), IDR-led GOPs of 30 pictures (1 I + 29 P,
CABAC-class syntax already entropy-decoded: header / motion-vector / coefficient buffers of
include/h264r.h), `--gops` independent GOPs per GPU.  A step reconstructs every picture of every GOP
on the rank's GPU: 30 dependency waves, each wave one batched launch per kernel family.  GOPs are the
sharding unit: every rank owns its own GOPs, there is no data-path collective; NCCL
only gathers the per-frame digests to rank 0 (weak scaling: per-GPU work is fixed).

  value  frames/s with the syntax buffers resident in HBM (device time of the kernels, CUDA events
         on the engine's stream, max over ranks)
  e2e    frames/s through the public C ABI with HOST buffers: pinned syntax -> H2D, kernels, digest
         D2H inside the timed region
  roofline / cpu_baseline: see DESIGN.md "Measurement"

`--impl reference` times the reference decoder itself (oracle/_ref/h264ref_harness = the reference's
own C++ compiled from /root/reference, queue-fed so that only its reconstruction path and macroblock
layer run) on the host cores, on a bounded sample of the same workload configuration.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W_MBS, H_MBS = 120, 68
N_MBS = W_MBS * H_MBS


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for k, n in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the compiled reference decoder on host cores
# --------------------------------------------------------------------------------------------------
def reference_sample(n_frames, seed=11):
    """Annex-B stream + syntax queue of one 1080p GOP prefix (1 I + n-1 P) with the bench content model."""
    from videodecoder_b200 import synth
    cache = os.path.join(tempfile.gettempdir(), "h264r_bench_sample_%d_%d" % (n_frames, seed))
    sp, qp = cache + ".264", cache + ".queue"
    if not (os.path.exists(sp) and os.path.exists(qp)):
        rng = np.random.default_rng(seed)
        seq = synth.make_seq(W_MBS, H_MBS)
        pics = synth.random_gop(rng, seq, n_frames, synth.ContentConfig(p_intra_in_p=0.03, p_skip=0.35, mvd_range=8))
        stream, queue = synth.serialize_stream(seq, pics)
        with open(sp + ".tmp", "wb") as f:
            f.write(stream)
        queue.tofile(qp + ".tmp")
        os.replace(sp + ".tmp", sp)
        os.replace(qp + ".tmp", qp)
    return sp, qp


def run_reference_pass(sp, qp, procs):
    """`procs` copies of the reference decoder in parallel, one per host core (GOP-parallel is the only
    parallelism the reference's design admits: it is single-threaded).  Returns (frames, wall seconds)."""
    harness = os.path.join(ROOT, "oracle", "_ref", "h264ref_harness")
    t0 = time.perf_counter()
    ps = [subprocess.Popen([harness, sp, qp, "-", "0"], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
          for _ in range(procs)]
    frames = 0
    for p in ps:
        _, err = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("reference harness failed: " + err[-500:])
        frames += json.loads(err.strip().splitlines()[-1])["pictures"]
    return frames, time.perf_counter() - t0


def reference_available():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "h264ref_harness"))


def bind_to_gpu_node(index):
    """CPU affinity of this process := the CPUs NVML reports as local to GPU `index`.  Returns a note for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return "host thread bound to the %d CPUs local to GPU %d" % (len(cpus), index)
    except Exception as e:      # no NVML, no permission: run unbound
        return "unbound (%s)" % type(e).__name__
    return "unbound"


def cpu_baseline(sample_frames=4):
    cores = os.cpu_count() or 1
    sp, qp = reference_sample(sample_frames)
    frames, secs = run_reference_pass(sp, qp, cores)
    return {"value": frames / secs, "unit": "frames/s", "cores": cores, "kind": "reference",
            "sample": "%d processes x %d pictures (1 I + %d P) of the 1080p workload, reference decoder compiled from its own "
                      "sources, queue-fed (no CABAC), wall %.1f s" % (cores, sample_frames, sample_frames - 1, secs)}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base = {"impl": "reference", "metric": "1080p_recon_frames_per_s", "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {"workload": "1080p 120x68 MBs, IDR-led GOP (1 I + P), REF_COMPAT syntax, bounded sample", "gop": args.sample_frames}}
    if not reference_available():
        base["unavailable"] = "oracle/_ref/h264ref_harness not built (no /root/reference on this box and no prebuilt binary)"
        print(json.dumps(base), flush=True)
        return
    cores = os.cpu_count() or 1
    sp, qp = reference_sample(args.sample_frames)
    for _ in range(args.warmup):
        run_reference_pass(sp, qp, cores)
    frames = 0
    secs = 0.0
    for _ in range(args.steps):
        f, s = run_reference_pass(sp, qp, cores)
        frames += f
        secs += s
    v = frames / secs
    base.update({"value": v, "ms_per_step": 1000.0 * secs / args.steps, "gpu_launches": 0,
                 "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": "reference",
                                  "sample": "%d processes x %d pictures per step" % (cores, args.sample_frames)},
                 "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(base), flush=True)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--gops", type=int, default=16, help="independent GOPs per GPU and step")
    ap.add_argument("--gop-len", type=int, default=30)
    ap.add_argument("--unique-gops", type=int, default=2, help="distinct GOP contents (replicated to --gops)")
    ap.add_argument("--mode", default="compat", choices=["compat", "spec"])
    ap.add_argument("--res", default="1080p", choices=["1080p", "4k"], help="picture size: 120x68 macroblocks (the metric's workload) or 240x135 (BASELINE config 5 geometry)")
    ap.add_argument("--sample-frames", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--layout", default="packed", choices=["packed", "dense"],
                    help="coefficient transport: h264r_submit_mbs_packed (present blocks only) or h264r_submit_mbs (768 B per macroblock)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl != "reference":
        log("note: fewer than 3 warm-up steps requested")
    if args.impl == "reference":
        reference_arm(args)
        return

    global W_MBS, H_MBS, N_MBS
    if args.res == "4k":
        W_MBS, H_MBS = 240, 135
        N_MBS = W_MBS * H_MBS
    import copy
    import torch
    import torch.distributed as dist
    from videodecoder_b200 import abi, engine, pack, workload

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path")
    torch.cuda.set_device(local)
    # The rank's host thread and its pinned picture buffers belong on the NUMA node of its GPU (what a decoding service
    # does with numactl): with 8 ranks the uploads otherwise cross the socket interconnect.  Restored before the CPU
    # baseline, which is to use every core.
    affinity_before = os.sched_getaffinity(0)
    numa_note = bind_to_gpu_node(local) if world > 1 and not os.environ.get("H264R_BENCH_NO_BIND") else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    compat = args.mode == "compat"
    flags = abi.REF_COMPAT if compat else 0
    G, F, U = args.gops, args.gop_len, min(args.unique_gops, args.gops)
    cfg = workload.WorkloadConfig(compat=compat, p_t8x8=0.0 if compat else 0.5, p_i8x8=0.0 if compat else 0.5,
                                  p_sub=(0.5, 0.0, 0.25, 0.25) if compat else (0.4, 0.2, 0.2, 0.2))
    t0 = time.time()
    rng = np.random.default_rng(1234)
    gops = [workload.make_gop(rng, W_MBS, H_MBS, F, cfg) for _ in range(U)]
    if compat:
        for g in gops:
            it = workload.sanitize(lambda n: engine.Engine(W_MBS, H_MBS, max_frames=n, flags=flags, device=local), W_MBS, H_MBS, g)
    log("rank %d: workload ready in %.1f s" % (rank, time.time() - t0))

    # digests of every frame are part of the job (GOP-sharded decoding compares them across ranks): K5 runs with each wave
    eng = engine.Engine(W_MBS, H_MBS, max_frames=G * F, flags=flags | abi.DIGESTS, device=local)
    # pinned host copies of the unique GOPs (what a host parser would hand over)
    packed = args.layout == "packed"
    coded_blocks = 0
    pinned = []
    inter_mbs_per_pic = []
    if packed:
        # one pinned picture buffer per picture of the step, in the engine's upload layout: what a host parser would
        # write into (the descriptor in front of each buffer is per frame, so buffers are not shared between GOPs)
        packed_src = [[pack.pack_blocks(p["coeff"]) for p in g] for g in gops]
        for g in range(G):
            frames = []
            for i, p in enumerate(gops[g % U]):
                mask, blocks = packed_src[g % U][i]
                pb = eng.picture_buf(max(1, blocks.shape[0]))
                pb.fill(p["hdr"], p["mv"], mask, blocks, mv_partition=True)
                if g < U:
                    coded_blocks += int(blocks.shape[0])
                frames.append((pb, None, None, None))
            pinned.append(frames)
    else:
        for g in gops:
            frames = []
            for p in g:
                hb = eng.pinned(np.uint8, (N_MBS, 16)); hb.array[:] = p["hdr"].view(np.uint8).reshape(N_MBS, 16)
                cb = eng.pinned(np.int16, (N_MBS, 384)); cb.array[:] = p["coeff"]
                mb = None
                if p["mv"] is not None:
                    mb = eng.pinned(np.int16, (N_MBS, 32)); mb.array[:] = p["mv"]
                frames.append((hb, mb, cb, None))
            pinned.append(frames)
    for g in gops:
        for p in g:
            inter_mbs_per_pic.append(int((p["hdr"]["mb_class"] >= abi.MB_P16x16).sum()) if p["mv"] is not None else 0)
    pps = []
    for g in range(G):
        for i in range(F):
            pp = copy.copy(gops[g % U][i]["pp"])
            pp2 = abi.PicParams.from_buffer_copy(pp)
            for k in range(pp2.num_ref_idx_l0):
                pp2.ref_list0[k] = g * F + gops[g % U][i]["pp"].ref_list0[k]
            pps.append(pp2)
    h2d = 0
    # bytes of one picture-buffer transfer: descriptor + headers + vectors + masks (fixed) + the blocks present
    upload_fixed = (pinned[0][0][0].c.blocks - pinned[0][0][0].c.mv_end) if packed else 0
    for g in range(G):
        for i in range(F):
            hb, mb, cb, kb = pinned[g if packed else g % U][i]
            if packed:
                h2d += upload_fixed + max(0, hb.n_mvs) * 4 + hb.n_blocks * {2: 32, 1: 16, 3: 8}[hb.c.level_bytes]
            else:
                h2d += hb.nbytes + cb.nbytes + (mb.nbytes if mb else 0)
    all_ids = list(range(G * F))

    def submit_all(flush_each=False):
        # decode order of a host that serves G streams round-robin: picture i of every GOP, then picture i + 1, ...
        # (uploads of later pictures overlap the kernels of earlier ones); flush_each: queue the kernels of every
        # such access-unit round at once instead of after the last submit
        L, ctx = eng.L, eng.ctx
        for i in range(F):
            if flush_each and i:
                eng.flush()
            for g in range(G):
                fid = g * F + i
                hb, mb, cb, kb = pinned[g if packed else g % U][i]
                rc = L.h264r_frame_begin(ctx, fid, pps[fid])
                if packed:
                    rc |= L.h264r_submit_picture(ctx, fid, hb.c, hb.n_blocks, hb.n_mvs)
                else:
                    rc |= L.h264r_submit_mbs(ctx, fid, 0, N_MBS, hb.ptr, mb.ptr if mb else None, cb.ptr)
                    rc |= L.h264r_frame_end(ctx, fid)
                if rc:
                    eng._check(rc)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def e2e_step():
        t0 = time.perf_counter()
        eng.mark(0)
        submit_all(flush_each=True)
        t1 = time.perf_counter()
        eng.flush()
        t2 = time.perf_counter()
        d = eng.digests(all_ids)
        eng.mark(1)
        ms = eng.elapsed_ms(0, 1)
        if os.environ.get("H264R_BENCH_DEBUG"):
            log("e2e step: submit %.2f ms, flush call %.2f ms, digests returned at %.2f ms, device %.2f ms" % (
                (t1 - t0) * 1e3, (t2 - t1) * 1e3, (time.perf_counter() - t0) * 1e3, ms))
        return ms, d

    for _ in range(args.warmup):
        _, dig0 = e2e_step()

    clocks = ClockSampler(local) if rank == 0 else None
    # ---- resident: kernels only, syntax already in HBM ------------------------------------------------
    barrier()
    torch.cuda.cudart().cudaProfilerStart()       # `ncu --profile-from-start off` sees exactly the timed resident steps
    kern_ms = 0.0
    for _ in range(args.steps):
        submit_all()
        eng.wait_uploads()              # inputs are resident in HBM before the timed flush
        eng.flush()
        ms, flush_launches = eng.last_flush_stats()       # events on the engine's stream around the whole flush
        kern_ms += ms
    torch.cuda.cudart().cudaProfilerStop()
    barrier()
    # ---- per-kernel times: the same steps on ONE lane with an event pair around every launch.  (In the timed steps
    # launches follow each other without events in between -- an event pair costs a short kernel tens of microseconds --
    # and, with H264R_LANES > 1, the kernels of the lanes overlap.)
    eng.set_lanes(1)
    eng.set_profiling(True)
    fam_ms = np.zeros(5)
    fam_n = np.zeros(5, dtype=np.int64)
    serial_ms = 0.0
    prof_steps = min(args.steps, 3)
    for _ in range(prof_steps):
        submit_all()
        eng.wait_uploads()
        eng.flush()
        serial_ms += eng.last_flush_stats()[0]
        fm, fn = eng.last_flush_kernel_ms()
        launch_ms = eng.last_flush_launch_ms()
        fam_ms += np.array(fm)
        fam_n += np.array(fn)
    eng.set_profiling(False)
    eng.set_lanes(int(os.environ.get("H264R_LANES", "1")))
    barrier()
    # ---- e2e: host buffers in, digests out -------------------------------------------------------------
    barrier()
    e2e_ms = 0.0
    for _ in range(args.steps):
        ms, dig = e2e_step()
        e2e_ms += ms
    barrier()
    # where the e2e time goes (untimed probe): host time of the submit calls, time until the uploads are resident
    t_a = time.perf_counter(); submit_all(); t_b = time.perf_counter(); eng.wait_uploads(); t_c = time.perf_counter()
    eng.flush(); eng.sync()
    e2e_probe = {"submit_calls_host_ms": (t_b - t_a) * 1e3, "uploads_resident_ms": (t_c - t_a) * 1e3,
                 "h2d_GBps_over_upload_window": h2d / max(1e-9, (t_c - t_a)) / 1e9}
    clk = clocks.stop() if clocks else None
    assert np.array_equal(dig, dig0), "digests changed between steps (non-deterministic reconstruction)"
    excursions = eng.range_excursions()

    t = torch.tensor([kern_ms, e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # the only collective of the system: per-frame digests to rank 0
        mine = torch.from_numpy(dig.copy()).cuda()
        gathered = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
        dist.gather(mine, gathered, dst=0)
        if rank == 0:
            for r, gd in enumerate(gathered):
                assert torch.equal(gd, gathered[0]), "rank %d reconstructed different pictures" % r
    kern_ms, e2e_ms = float(t[0]), float(t[1])
    frames_per_step = G * F * world

    if rank == 0:
        peak, peak_src = measured_peak()
        # dominant kernel family and its roofline (DESIGN.md "Measurement")
        fam = int(np.argmax(fam_ms[:3]))
        names = ["k_inter (K1+K3: dequant/IDCT + motion compensation + reconstruction)",
                 "k_intra (K1+K2: dequant/IDCT + intra wavefront)", "k_deblock (K4)"]
        inter_mbs = sum(inter_mbs_per_pic[(g % U) * F + i] for g in range(G) for i in range(F))
        per_mb = {0: (768 + 16 + 64 + (256 if compat else 384) + 384), 1: (768 + 16 + 384), 2: (384 + 384 + 16 + 64)}
        if fam == 0:
            units = inter_mbs
        elif fam == 1:
            units = G * F * N_MBS - inter_mbs
        else:
            units = G * F * N_MBS
        bytes_per_launch = per_mb[fam] * units / max(1, fam_n[fam] / prof_steps)
        sec_per_launch = fam_ms[fam] / 1000.0 / max(1, fam_n[fam])
        achieved = bytes_per_launch / sec_per_launch / 1e9 if sec_per_launch > 0 else 0.0
        # DRAM bytes per launch of that kernel from the committed ncu --set full capture of this very workload (null when
        # the launch size differs from the captured one)
        traffic = None
        tj = os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")
        if os.path.exists(tj):
            try:
                tr = json.load(open(tj))
                if tr.get("_pictures_per_launch") == G:
                    traffic = tr[args.mode][["k_inter", "k_intra_rows", "k_deblock_rows"][fam]]["dram_bytes_per_launch"]
            except Exception:
                traffic = None
        line = {
            "metric": "1080p_recon_frames_per_s" if args.res == "1080p" else "4k_recon_frames_per_s", "value": frames_per_step * args.steps / (kern_ms / 1000.0), "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": kern_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": "%s %dx%d MBs, %d GOPs/GPU x %d pictures (1 I + %d P), %s" % (
                args.res, W_MBS, H_MBS, G, F, F - 1, "REF_COMPAT (what the reference decoder computes: 4x4 transform, luma prediction, no deblock)" if compat
                else "spec mode (chroma prediction + MC, 8x8 transform, Intra8x8, deblocking)"),
                "mode": args.mode, "layout": args.layout + (" (%.2f of 24 coefficient blocks per macroblock present; int8 levels in 8-byte units: significance map + levels)" % (coded_blocks / float(U * F * N_MBS)) if packed else ""),
                "gops_per_gpu": G, "gop_len": F, "unique_gops": U, "sharding": "one GOP set per rank, no data-path collective",
                "l2_policy": "inputs (%.1f GB of syntax per step) exceed the 126 MB L2 and are re-uploaded between timed steps" % (h2d / 1e9)},
            "e2e": {"value": frames_per_step * args.steps / (e2e_ms / 1000.0), "unit": "frames/s", "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": G * F * 16},
            # kernels launched inside the timed regions: the resident steps (one flush each) + the e2e steps (the same
            # kernels, queued round by round, + one digest launch)
            "gpu_launches": int(args.steps * (2 * flush_launches + 1)),
            "kernel_ms_per_step": {"inter": fam_ms[0] / prof_steps, "intra": fam_ms[1] / prof_steps, "deblock": fam_ms[2] / prof_steps, "pad": fam_ms[3] / prof_steps, "digest": fam_ms[4] / prof_steps,
                                   "step_on_one_lane": serial_ms / prof_steps,
                                   "note": "per-launch events, one lane, %d extra steps after the timed ones; event pairs inflate the short launches" % prof_steps},
            "roofline": {"bound": "hbm", "kernel": names[fam], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src, "traffic": traffic, "algorithmic_bytes_per_launch": bytes_per_launch,
                         "avg_launch_ms": sec_per_launch * 1000.0},
            "clocks": clk, "luma_range_excursions": excursions,
        }
        # per-launch picture of the last timed flush: the first wave holds the I pictures, the others the P pictures
        prof = {}
        for fam_id, fam_name in ((0, "inter"), (1, "intra"), (2, "deblock"), (3, "pad"), (4, "digest")):
            t = [m for m, f in launch_ms if f == fam_id]
            if t:
                prof[fam_name] = {"n": len(t), "first_ms": t[0], "rest_avg_ms": (sum(t[1:]) / (len(t) - 1)) if len(t) > 1 else None,
                                  "max_ms": max(t)}
        line["launch_profile"] = prof
        line["e2e_probe"] = e2e_probe
        if numa_note:
            line["config"]["host_binding"] = numa_note
        os.sched_setaffinity(0, affinity_before)
        if world == 1 and not args.no_cpu_baseline and reference_available():
            try:
                line["cpu_baseline"] = cpu_baseline(args.sample_frames)
            except Exception as e:  # the baseline must not lose the measurement
                line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %s" % e}
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
