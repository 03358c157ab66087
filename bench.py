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

  configs   (N = 1) the other single-GPU configurations of BASELINE.json, each with its own value / e2e / roofline:
            config 3 (1080p High: 8x8 transform, Intra8x8, chroma prediction + MC, deblocking) and config 5's geometry (4K)
  e2e_planes  like e2e, but every reconstructed frame also travels back to pinned host memory (h264r_read_frame_async)
  host      the serial host side (libh264host.so: Annex-B + CABAC -> picture buffers) on a genuine bitstream of the same
            content, frames/s per core -- what a producer of the e2e buffers costs
  config4   (N > 1) BASELINE config 4 as written: 20 distinct GOPs, GOP g on rank g mod N, per-frame digests gathered on rank 0
            and compared with the 1-GPU list committed under profiles/

`--impl reference` times the reference decoder itself (oracle/_ref/h264ref_harness = the reference's
own C++ compiled from /root/reference, queue-fed so that only its reconstruction path and macroblock
layer run) on the host cores, on a bounded sample of THE SAME workload: the first pictures of GOP 0 of the generator
and seed the GPU arm uses, turned into the reference's syntax by pack.syntax_from_buffers + the stream writer.
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
WORKLOAD_SEED = 1234


def workload_cfg(compat):
    from videodecoder_b200 import workload
    return workload.WorkloadConfig(compat=compat, p_t8x8=0.0 if compat else 0.5, p_i8x8=0.0 if compat else 0.5,
                                   p_sub=(0.5, 0.0, 0.25, 0.25) if compat else (0.4, 0.2, 0.2, 0.2))


def config_of(args):
    """`config` of the JSON line: what the workload IS (derived from the arguments only, so that both arms print the same)."""
    compat = args.mode == "compat"
    w, h = (240, 135) if args.res == "4k" else (120, 68)
    return {"workload": "%s %dx%d MBs, IDR-led GOPs of %d pictures (1 I + %d P), %d GOPs per GPU and step, %s; videodecoder_b200/workload.py content model, seed %d" % (
                args.res, w, h, args.gop_len, args.gop_len - 1, args.gops,
                "REF_COMPAT (what the reference decoder computes: 4x4 transform, luma prediction, no deblock)" if compat
                else "spec mode (chroma prediction + MC, 8x8 transform, Intra8x8, deblocking)", WORKLOAD_SEED),
            "mode": args.mode, "layout": args.layout, "headers": ({"stream": "stream form (8-byte headers + one stream per picture, H264R_HDR_STREAM)", "short": "8-byte (h264r_mb_hdr8)", "long": "16-byte"}[args.hdr] if args.layout == "packed" else "16-byte"), "gops_per_gpu": args.gops, "gop_len": args.gop_len, "unique_gops": min(args.unique_gops, args.gops),
            "sharding": "one GOP set per rank, no data-path collective",
            "l2_policy": "the syntax of a step exceeds the 126 MB L2 and is re-uploaded between timed steps"}


def reference_sample(n_frames, sub8x8_only=False):
    """The first n_frames pictures of GOP 0 of the benchmark workload (same generator, same seed, REF_COMPAT content kept inside
    [0,255] the same way -- here through the CPU oracle's excursion map, on the GPU arm through the engine's), as the reference
    decoder reads them: Annex-B stream + syntax queue for the queue-fed harness, and a genuine CABAC stream for the
    host-decoder measurement.  sub8x8_only: sub-macroblock partitions switched off (the reference's own CABAC cannot read
    them).  Returns (stream path, queue path, cabac stream path)."""
    from videodecoder_b200 import abi, hoststream, pack, synth, workload
    cache = os.path.join(tempfile.gettempdir(), "h264r_bench_ref_%d_%d_%d" % (n_frames, WORKLOAD_SEED, int(sub8x8_only)))
    sp, qp, cp = cache + ".264", cache + ".queue", cache + ".cabac.264"
    if not (os.path.exists(sp) and os.path.exists(qp) and os.path.exists(cp)):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle                      # the checker, used here to prepare the reference leg's input
        cfg = workload_cfg(True)
        if sub8x8_only:
            cfg.p_sub = (1.0, 0.0, 0.0, 0.0)
        rng = np.random.default_rng(WORKLOAD_SEED)
        gop = [workload.make_picture(rng, W_MBS, H_MBS, i > 0, cfg, [i - 1] if i else []) for i in range(n_frames)]
        outs = []
        for p in gop:                      # same rule as workload.sanitize: zero the residual of macroblocks that leave [0,255]
            pp = p["pp"]
            refs = [(outs[pp.ref_list0[i]]["y"], outs[pp.ref_list0[i]]["u"], outs[pp.ref_list0[i]]["v"]) for i in range(pp.num_ref_idx_l0)]
            for _ in range(24):
                o = oracle.recon_picture(W_MBS, H_MBS, abi.REF_COMPAT, pp, p["hdr"], p["mv"], p["coeff"], refs=refs)
                if o["excursions"] == 0:
                    break
                workload.zero_mbs(p, np.flatnonzero(o["excursion_map"]))
            outs.append(o)
        seq = synth.make_seq(W_MBS, H_MBS)
        pics = [(synth.make_pic(seq, i, synth.SLICE_P if i else synth.SLICE_I), pack.syntax_from_buffers(p["hdr"], p["mv"], p["coeff"], seq.pic_init_qp))
                for i, p in enumerate(gop)]
        cabac = hoststream.encode_stream(seq, pics, flags=hoststream.E_ABSOLUTE | hoststream.E_REF_QUIRKS)     # also rewrites mvd / prev / rem in place
        stream, queue = synth.serialize_stream(seq, pics)
        for path, data in ((sp, stream), (cp, cabac + b"\x00\x00\x00\x01\x0c\xff\x80")):
            with open(path + ".tmp", "wb") as f:
                f.write(data)
            os.replace(path + ".tmp", path)
        queue.tofile(qp + ".tmp")
        os.replace(qp + ".tmp", qp)
    return sp, qp, cp


def run_reference_pass(sp, qp, procs, own_cabac=False):
    """`procs` copies of the reference decoder in parallel, one per host core (GOP-parallel is the only
    parallelism the reference's design admits: it is single-threaded).  Returns (frames, wall seconds).
    own_cabac: the reference with its own arithmetic decoder on a genuine CABAC stream (sp) instead of the queue-fed one."""
    harness = os.path.join(ROOT, "oracle", "_ref", "h264ref_cabac" if own_cabac else "h264ref_harness")
    t0 = time.perf_counter()
    ps = [subprocess.Popen([harness, sp, "-" if own_cabac else qp, "-", "0"], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
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


SAMPLE_NOTE = ("%d processes x the first %d pictures (1 I + %d P) of GOP 0 of this workload (same generator and seed as the GPU arm), "
               "reference decoder compiled from its own sources, queue-fed (its macroblock layer + reconstruction, no CABAC), wall %.1f s")


def cpu_baseline(sample_frames=4):
    cores = os.cpu_count() or 1
    sp, qp, _ = reference_sample(sample_frames)
    frames, secs = run_reference_pass(sp, qp, cores)
    out = {"value": frames / secs, "unit": "frames/s", "cores": cores, "kind": "reference",
           "sample": SAMPLE_NOTE % (cores, sample_frames, sample_frames - 1, secs)}
    # the whole reference decoder (its own bit reader + CABAC too) on a genuine bitstream, for scale; needs the variant of the
    # sample without sub-8x8 partitions (its CABAC cannot read them)
    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "h264ref_cabac")):
        try:
            _, _, cp = reference_sample(sample_frames, sub8x8_only=True)
            f2, s2 = run_reference_pass(cp, None, cores, own_cabac=True)
            out["full_decoder"] = {"value": f2 / s2, "unit": "frames/s", "cores": cores,
                                   "note": "unmodified reference incl. its own CABAC on a genuine bitstream of the same sample with 8x8 sub-partitions only"}
        except Exception as e:
            out["full_decoder"] = {"value": None, "note": "failed: %s" % e}
    return out


def host_side(sample_frames=4):
    """The serial host side of THIS repository (libh264host.so) on a genuine CABAC bitstream of the sample: frames/s on one core."""
    from videodecoder_b200 import hoststream
    _, _, cp = reference_sample(sample_frames)
    data = open(cp, "rb").read()
    hoststream.benchmark(data, flags=hoststream.REF_QUIRKS, repeat=1)
    n, sec = hoststream.benchmark(data, flags=hoststream.REF_QUIRKS, repeat=3)
    n2, sec2 = hoststream.benchmark(data, flags=hoststream.REF_QUIRKS | hoststream.BENCH_PACK, repeat=3)
    return {"parse_frames_per_s_per_core": n / sec, "parse_and_pack_frames_per_s_per_core": n2 / sec2,
            "stream_bytes_per_frame": len(data) / sample_frames,
            "what": "libh264host.so (Annex-B, CABAC, derivations, DPB -> dense per-macroblock arrays; + h264r_writer.h into the stream form of a picture buffer) on a CABAC stream of the first %d pictures of GOP 0, one thread" % sample_frames}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base = {"impl": "reference", "metric": "1080p_recon_frames_per_s", "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic", "config": config_of(args)}
    if not reference_available():
        base["unavailable"] = "oracle/_ref/h264ref_harness not built (no /root/reference on this box and no prebuilt binary)"
        print(json.dumps(base), flush=True)
        return
    cores = os.cpu_count() or 1
    sp, qp, _ = reference_sample(args.sample_frames)
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
                                  "sample": SAMPLE_NOTE % (cores, args.sample_frames, args.sample_frames - 1, secs / max(1, args.steps))},
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


class Cfg:
    """one measured configuration (what argparse gives for the headline, built by hand for the `configs` block)"""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def measure(a, rank, world, local, dist, torch, extras=True):
    """Runs one configuration: resident steps (value), per-kernel profile (roofline), end-to-end steps (e2e).  Returns the
    fields of the JSON line that depend on the configuration (rank 0; other ranks return None)."""
    import copy
    from videodecoder_b200 import abi, engine, pack, workload
    Wm, Hm = (240, 135) if a.res == "4k" else (120, 68)
    n_mbs = Wm * Hm
    compat = a.mode == "compat"
    flags = abi.REF_COMPAT if compat else 0
    G, F, U = a.gops, a.gop_len, min(a.unique_gops, a.gops)
    cfg = workload_cfg(compat)
    if a.p_block is not None:
        cfg.p_block = a.p_block
    if a.max_amp is not None:
        cfg.max_amp = a.max_amp
    t0 = time.time()
    rng = np.random.default_rng(WORKLOAD_SEED)
    gops = [workload.make_gop(rng, Wm, Hm, F, cfg) for _ in range(U)]
    if compat:
        for g in gops:
            workload.sanitize(lambda n: engine.Engine(Wm, Hm, max_frames=n, flags=flags, device=local), Wm, Hm, g)
    log("rank %d: %s %s workload ready in %.1f s" % (rank, a.res, a.mode, time.time() - t0))

    # digests of every frame are part of the job (GOP-sharded decoding compares them across ranks): K5 runs with each wave
    eng = engine.Engine(Wm, Hm, max_frames=G * F, flags=flags | abi.DIGESTS, device=local)
    packed = a.layout == "packed"
    rings = None
    coded_blocks = 0
    level_stats = [0, 0]
    pinned = []
    inter_mbs_per_pic = []
    if packed:
        # one pinned picture buffer per picture of the step, in the engine's upload layout: what a host parser writes into
        packed_src = [[pack.pack_blocks(p["coeff"]) for p in g] for g in gops]
        # one RING of picture buffers per round of access units (picture i of every GOP): a round travels as one strided DMA
        rings = [eng.picture_ring(G, max(1, max(packed_src[u][i][1].shape[0] for u in range(U))) + 2 * n_mbs) for i in range(F)] if a.ring else None
        for g in range(G):
            frames = []
            for i, p in enumerate(gops[g % U]):
                mask, blocks = packed_src[g % U][i]
                # (+ 2 blocks' worth of bytes per macroblock: in the stream form the vectors and masks share the block area)
                pb = rings[i][g] if rings else eng.picture_buf(max(1, blocks.shape[0]) + (2 * n_mbs if a.hdr == "stream" else 0))
                if not (a.hdr == "stream" and pb.fill_stream(p["hdr"], p["mv"], p["coeff"])):
                    pb.fill(p["hdr"], p["mv"], mask, blocks, mv_partition=True, short_hdr=a.hdr != "long")
                if g < U:
                    coded_blocks += int(blocks.shape[0])
                    level_stats[0] += int((blocks != 0).sum())
                    level_stats[1] += int(pb.c.level_bytes == abi.LEVELS_COMPACT)
                frames.append((pb, None, None, None))
            pinned.append(frames)
    else:
        for g in gops:
            frames = []
            for p in g:
                hb = eng.pinned(np.uint8, (n_mbs, 16)); hb.array[:] = p["hdr"].view(np.uint8).reshape(n_mbs, 16)
                cb = eng.pinned(np.int16, (n_mbs, 384)); cb.array[:] = p["coeff"]
                mb = None
                if p["mv"] is not None:
                    mb = eng.pinned(np.int16, (n_mbs, 32)); mb.array[:] = p["mv"]
                frames.append((hb, mb, cb, None))
            pinned.append(frames)
    for g in gops:
        for p in g:
            inter_mbs_per_pic.append(int((p["hdr"]["mb_class"] >= abi.MB_P16x16).sum()) if p["mv"] is not None else 0)
    pps = []
    for g in range(G):
        for i in range(F):
            pp2 = abi.PicParams.from_buffer_copy(copy.copy(gops[g % U][i]["pp"]))
            for k in range(pp2.num_ref_idx_l0):
                pp2.ref_list0[k] = g * F + gops[g % U][i]["pp"].ref_list0[k]
            pps.append(pp2)
    h2d = 0
    # bytes of the transfers: a picture buffer travels from its first vector (they grow backward from the descriptor) to its last
    # unit; a ring round as one strided DMA over the union of its pictures' ranges
    def span(hb):
        base = hb.c.base
        if hb.c.hdr_bytes == abi.HDR_STREAM:          # descriptor, 8-byte headers, word counts, the stream
            return hb.c.mv_end - base, (hb.c.stream - base) + hb.n_blocks * 4
        return (hb.c.mv_end - base) - max(0, hb.n_mvs) * 4, (hb.c.blocks - base) + hb.n_blocks * {2: 32, 1: 16, 3: 8}[hb.c.level_bytes]
    for i in range(F):
        spans = []
        for g in range(G):
            hb, mb, cb, kb = pinned[g if packed else g % U][i]
            if packed:
                spans.append(span(hb))
            else:
                h2d += hb.nbytes + cb.nbytes + (mb.nbytes if mb else 0)
        if packed and rings:
            h2d += (max(e for _, e in spans) - min(f for f, _ in spans)) * G
        elif packed:
            h2d += sum(e - f for f, e in spans)
    all_ids = list(range(G * F))

    def submit_round(i):
        L, ctx = eng.L, eng.ctx
        if packed and rings:
            eng.submit_ring(rings[i], [g * F + i for g in range(G)], [pps[g * F + i] for g in range(G)])
            return
        for g in range(G):
            fid = g * F + i
            hb, mb, cb, kb = pinned[g if packed else g % U][i]
            rc = L.h264r_frame_begin(ctx, fid, pps[fid])
            if packed:
                rc |= L.h264r_submit_picture(ctx, fid, hb.c, hb.n_blocks, hb.n_mvs)
            else:
                rc |= L.h264r_submit_mbs(ctx, fid, 0, n_mbs, hb.ptr, mb.ptr if mb else None, cb.ptr)
                rc |= L.h264r_frame_end(ctx, fid)
            if rc:
                eng._check(rc)

    def submit_all(flush_each=False):
        # decode order of a host that serves G streams round-robin: picture i of every GOP, then picture i + 1, ...
        # (uploads of later pictures overlap the kernels of earlier ones); flush_each: queue the kernels of every
        # such access-unit round at once instead of after the last submit
        for i in range(F):
            if flush_each and i:
                eng.flush()
            submit_round(i)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def e2e_step():
        eng.mark(0)
        submit_all(flush_each=True)
        eng.flush()
        d = eng.digests(all_ids)
        eng.mark(1)
        return eng.elapsed_ms(0, 1), d

    for _ in range(a.warmup):
        _, dig0 = e2e_step()

    clocks = ClockSampler(local) if rank == 0 else None
    # ---- resident: kernels only, syntax already in HBM ------------------------------------------------
    barrier()
    torch.cuda.cudart().cudaProfilerStart()       # `ncu --profile-from-start off` sees exactly the timed resident steps
    kern_ms = 0.0
    for _ in range(a.steps):
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
    prof_steps = min(a.steps, 3)
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
    e2e_step()                          # untimed: the per-kernel section above ran on one lane with event pairs; settle back first
    barrier()
    e2e_list = []
    for _ in range(a.steps):
        ms, dig = e2e_step()
        e2e_list.append(ms)
    barrier()
    e2e_ms = float(sum(e2e_list))
    # where the e2e time goes (untimed probe): host time of the submit calls, time until the uploads are resident
    t_a = time.perf_counter(); submit_all(); t_b = time.perf_counter(); eng.wait_uploads(); t_c = time.perf_counter()
    eng.flush(); eng.sync()
    # host time of the e2e loop itself (submits + a flush per round, nothing waited for): if it approaches the e2e step time, the
    # host thread that issues the work bounds the number, not the GPU
    t_d = time.perf_counter(); submit_all(flush_each=True); eng.flush(); t_e = time.perf_counter()
    eng.sync()
    e2e_probe = {"submit_calls_host_ms": (t_b - t_a) * 1e3, "uploads_resident_ms": (t_c - t_a) * 1e3,
                 "h2d_GBps_over_upload_window": h2d / max(1e-9, (t_c - t_a)) / 1e9,
                 "host_issue_ms_submit_and_flush_per_round": (t_e - t_d) * 1e3}
    # ---- e2e_planes: as e2e, and every reconstructed frame goes back to pinned host memory -------------------------------
    planes = None
    if extras and a.planes:
        lay = eng.frame_layout()
        fb = int(lay.frame_bytes)
        try:
            host = eng.pinned(np.uint8, (G * F * fb,))
        except Exception as e:          # not enough pinned memory on this box: say so instead of failing the run
            host = None
            planes = {"value": None, "note": "pinned allocation of %.1f GB failed: %s" % (G * F * fb / 1e9, e)}
        if host is not None:
            def planes_step():
                eng.mark(2)
                for i in range(F):
                    submit_round(i)
                    eng.flush()
                    for g in range(G):      # the frames of this round travel back while the next rounds are reconstructed
                        eng.read_frame_async(g * F + i, host.ptr + (g * F + i) * fb)
                d = eng.digests(all_ids)
                eng.wait_readbacks()
                eng.mark(3)
                return eng.elapsed_ms(2, 3), d
            planes_step()
            pl = []
            for _ in range(min(a.steps, 3)):
                ms, dp = planes_step()
                pl.append(ms)
            assert np.array_equal(dp, dig0)
            # the frames that arrived are the frames the device holds: spot check against h264r_read_planes
            for fid in (0, G * F - 1):
                y0, u0, v0 = eng.read_planes(fid)
                y1, u1, v1 = engine.Engine.planes_of(lay, host.array[fid * fb:(fid + 1) * fb])
                assert np.array_equal(y0, y1) and np.array_equal(u0, u1) and np.array_equal(v0, v1), "read-back frame differs"
            t = torch.tensor([float(np.median(pl))], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            planes = {"value": G * F * world / (float(t[0]) / 1000.0), "unit": "frames/s", "ms_per_step_median": float(t[0]),
                      "d2h_bytes_per_step": G * F * fb + G * F * 16, "h2d_bytes_per_step": h2d,
                      "note": "every frame slot (planes + replicated border, %d bytes) copied to pinned host memory on a read-back stream, overlapped with the reconstruction of later rounds" % fb}
            host.free()
    clk = clocks.stop() if clocks else None
    assert np.array_equal(dig, dig0), "digests changed between steps (non-deterministic reconstruction)"
    excursions = eng.range_excursions()

    t = torch.tensor([kern_ms, e2e_ms, float(np.median(e2e_list)), float(min(e2e_list))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # the only collective of the system: per-frame digests to rank 0
        mine = torch.from_numpy(dig.copy()).cuda()
        gathered = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
        dist.gather(mine, gathered, dst=0)
        if rank == 0:
            for r, gd in enumerate(gathered):
                assert torch.equal(gd, gathered[0]), "rank %d reconstructed different pictures" % r
    kern_ms, e2e_ms, e2e_med, e2e_best = (float(x) for x in t)
    frames_per_step = G * F * world
    out = None
    if rank == 0:
        peak, peak_src = measured_peak()
        # dominant kernel family and its roofline (DESIGN.md "Measurement")
        fam = int(np.argmax(fam_ms[:3]))
        names = [("k_inter_grp" if (compat and packed) else "k_inter") + " (K1+K3: dequant/IDCT + motion compensation + reconstruction)",
                 "k_intra (K1+K2: dequant/IDCT + intra wavefront)", "k_deblock (K4)"]
        inter_mbs = sum(inter_mbs_per_pic[(g % U) * F + i] for g in range(G) for i in range(F))
        per_mb = {0: (768 + 16 + 64 + (256 if compat else 384) + 384), 1: (768 + 16 + 384), 2: (384 + 384 + 16 + 64)}
        units = inter_mbs if fam == 0 else (G * F * n_mbs - inter_mbs if fam == 1 else G * F * n_mbs)
        bytes_per_launch = per_mb[fam] * units / max(1, fam_n[fam] / prof_steps)
        sec_per_launch = fam_ms[fam] / 1000.0 / max(1, fam_n[fam])
        achieved = bytes_per_launch / sec_per_launch / 1e9 if sec_per_launch > 0 else 0.0
        # DRAM bytes per launch of that kernel from the committed ncu --set full capture of this very workload (null when
        # the launch size differs from the captured one)
        traffic = None
        for tj in ("r2_ncu_traffic.json", "r1_ncu_traffic.json"):
            tj = os.path.join(ROOT, "profiles", tj)
            if os.path.exists(tj) and a.res == "1080p":
                try:
                    tr = json.load(open(tj))
                    if tr.get("_pictures_per_launch") == G:
                        # (REF_COMPAT launches of picture buffers run k_inter_grp; the capture lists the kernel that ran)
                        fam_kernels = [["k_inter_grp", "k_inter"] if a.mode == "compat" else ["k_inter"], ["k_intra_rows"], ["k_deblock_rows"]][fam]
                        hit = [k for k in fam_kernels if k in tr[a.mode]]
                        if hit:
                            traffic = tr[a.mode][hit[0]]["dram_bytes_per_launch"]
                            break
                except Exception:
                    traffic = None
        out = {
            "value": frames_per_step * a.steps / (kern_ms / 1000.0), "ms_per_step": kern_ms / a.steps,
            "e2e": {"value": frames_per_step * a.steps / (e2e_ms / 1000.0), "unit": "frames/s", "ms_per_step": e2e_ms / a.steps,
                    "median_of_steps": frames_per_step / (e2e_med / 1000.0), "best_of_steps": frames_per_step / (e2e_best / 1000.0),
                    "steps_ms": [round(x, 3) for x in e2e_list],
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": G * F * 16},
            # kernels launched inside the timed regions: the resident steps (one flush each) + the e2e steps (the same
            # kernels, queued round by round, + one digest launch)
            "gpu_launches": int(a.steps * (2 * flush_launches + 1)),
            "kernel_ms_per_step": {"inter": fam_ms[0] / prof_steps, "intra": fam_ms[1] / prof_steps, "deblock": fam_ms[2] / prof_steps, "pad": fam_ms[3] / prof_steps, "digest": fam_ms[4] / prof_steps,
                                   "step_on_one_lane": serial_ms / prof_steps,
                                   "note": "per-launch events, one lane, %d extra steps after the timed ones; event pairs inflate the short launches" % prof_steps},
            "roofline": {"bound": "hbm", "kernel": names[fam], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src, "traffic": traffic,
                         "frac_on_dram_traffic": (traffic / sec_per_launch / 1e9 / peak) if traffic and sec_per_launch > 0 else None,
                         "algorithmic_bytes_per_launch": bytes_per_launch, "avg_launch_ms": sec_per_launch * 1000.0},
            "clocks": clk, "luma_range_excursions": excursions,
            "content": {"coded_blocks_per_mb": coded_blocks / float(U * F * n_mbs) if packed else None,
                        "levels_per_coded_block": (level_stats[0] / max(1, coded_blocks)) if packed else None,
                        "pictures_in_compact_transport": (level_stats[1] / float(U * F)) if packed else None,
                        "h2d_bytes_per_mb": h2d / float(G * F * n_mbs), "p_block": cfg.p_block, "max_amp": cfg.max_amp},
        }
        # per-launch picture of the last timed flush: the first wave holds the I pictures, the others the P pictures
        prof = {}
        for fam_id, fam_name in ((0, "inter"), (1, "intra"), (2, "deblock"), (3, "pad"), (4, "digest")):
            tt = [m for m, f in launch_ms if f == fam_id]
            if tt:
                prof[fam_name] = {"n": len(tt), "first_ms": tt[0], "rest_avg_ms": (sum(tt[1:]) / (len(tt) - 1)) if len(tt) > 1 else None,
                                  "max_ms": max(tt)}
        out["launch_profile"] = prof
        out["e2e_probe"] = e2e_probe
        if planes is not None:
            out["e2e_planes"] = planes
    for g in pinned:
        for fr in g:
            for b in fr:
                if b is not None:
                    b.free()
    eng.close()
    return out


def config4(rank, world, local, dist, torch):
    """BASELINE config 4 as written: 1080p, 600 pictures = 20 IDR-delimited GOPs x 30, GOP g reconstructed by rank g mod N
    (videodecoder_b200/sharding.py), no data-path collective; the 16-byte digests of all 600 frames are gathered on rank 0 and
    compared with the list a single GPU produced (profiles/r2_config4_digests.json).  REF_COMPAT syntax straight from the
    generator (content is not range-restricted here: the check is determinism across the sharding, not parity)."""
    import hashlib
    from videodecoder_b200 import abi, engine, pack, sharding, workload
    n_gops, F = 20, 30
    mine = sharding.assign_gops(n_gops, world)[rank]
    cfg = workload_cfg(True)
    eng = engine.Engine(W_MBS, H_MBS, max_frames=max(1, len(mine)) * F, flags=abi.REF_COMPAT | abi.DIGESTS, device=local)
    bufs = []
    t_gen = time.time()
    for k, g in enumerate(mine):
        gop = workload.make_gop(np.random.default_rng(4000 + g), W_MBS, H_MBS, F, cfg)
        row = []
        for i, p in enumerate(gop):
            mask, blocks = pack.pack_blocks(p["coeff"])
            pb = eng.picture_buf(max(1, blocks.shape[0]) + 2 * W_MBS * H_MBS)
            if not pb.fill_stream(p["hdr"], p["mv"], p["coeff"]):
                pb.fill(p["hdr"], p["mv"], mask, blocks, mv_partition=True, short_hdr=True)
            pp = abi.PicParams.from_buffer_copy(p["pp"])
            for r in range(pp.num_ref_idx_l0):
                pp.ref_list0[r] = k * F + p["pp"].ref_list0[r]
            row.append((pb, pp))
        bufs.append(row)
    t_gen = time.time() - t_gen

    def step():
        eng.mark(0)
        for i in range(F):
            if i:
                eng.flush()
            for k in range(len(mine)):
                pb, pp = bufs[k][i]
                eng.submit_picture_buf(k * F + i, pp, pb)
        eng.flush()
        d = eng.digests(list(range(len(mine) * F))) if mine else np.zeros((0, 16), np.uint8)
        eng.mark(1)
        return eng.elapsed_ms(0, 1), d
    step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms, dig = step()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # the only collective of the system: the per-frame digests to rank 0, put back in GOP order (sharding.gather_digests)
    table = sharding.gather_digests(dig, mine, n_gops, F, rank, world, device="cuda")
    out = None
    if rank == 0:
        md5 = hashlib.md5(np.ascontiguousarray(table).tobytes()).hexdigest()
        want = None
        ref = os.path.join(ROOT, "profiles", "r2_config4_digests.json")
        if os.path.exists(ref):
            want = json.load(open(ref)).get("md5_of_600_digests")
        out = {"frames": n_gops * F, "gops": n_gops, "sharding": "GOP g -> rank g mod %d" % world, "ms": float(t[0]),
               "value": n_gops * F / (float(t[0]) / 1000.0), "unit": "frames/s", "scaling": "strong",
               "md5_of_600_digests": md5, "matches_1gpu_list": (md5 == want) if want else None,
               "workload_generation_s_per_rank": t_gen}
    for row in bufs:
        for pb, _ in row:
            pb.free()
    eng.close()
    return out


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
    ap.add_argument("--no-extras", action="store_true", help="headline configuration only (no configs block, no e2e_planes, no host-side figure)")
    ap.add_argument("--config4", action="store_true", help="also run BASELINE config 4 (20 distinct GOPs, GOP g on rank g mod N); default at N > 1")
    ap.add_argument("--layout", default="packed", choices=["packed", "dense"],
                    help="coefficient transport: h264r_submit_mbs_packed (present blocks only) or h264r_submit_mbs (768 B per macroblock)")
    ap.add_argument("--hdr", default="stream", choices=["stream", "short", "long"],
                    help="picture buffers in stream form (8-byte headers + one stream per picture: H264R_HDR_STREAM), with 8-byte (h264r_mb_hdr8) or with 16-byte headers")
    ap.add_argument("--ring", action="store_true", help="picture rings: one strided DMA per round of access units instead of one DMA per picture "
                    "(h264r_submit_pictures; 52 against 33 GB/s on one GPU, but the union of the pictures' byte ranges travels -- +8 %% bytes here, "
                    "which costs 5 %% at 8 GPUs, where the box's host memory bounds the uploads)")
    ap.add_argument("--p-block", type=float, default=None, help="density sweep: share of coded 4x4 blocks inside a coded 8x8 (default 0.45)")
    ap.add_argument("--max-amp", type=float, default=None, help="density sweep: residual amplitude (default 10)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl != "reference":
        log("note: fewer than 3 warm-up steps requested")
    if args.impl == "reference":
        reference_arm(args)
        return

    import torch
    import torch.distributed as dist

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
    extras = world == 1 and not args.no_extras and args.p_block is None and args.max_amp is None and args.layout == "packed"
    head = Cfg(**vars(args))
    head.planes = True
    m = measure(head, rank, world, local, dist, torch, extras=extras or world > 1)
    line = None
    if rank == 0:
        line = {"metric": "1080p_recon_frames_per_s" if args.res == "1080p" else "4k_recon_frames_per_s", "value": m.pop("value"), "unit": "frames/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": m.pop("ms_per_step"),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                "config": config_of(args)}
        line.update(m)
        if numa_note:
            line["config"]["host_binding"] = numa_note
    if extras and args.mode == "compat" and args.res == "1080p":
        # the other single-GPU configurations of BASELINE.json, shorter runs of the same procedure
        others = {"config3_1080p_high_spec_mode": Cfg(**{**vars(args), "mode": "spec", "res": "1080p", "steps": 3, "warmup": 3, "planes": False}),
                  "config5_4k_geometry": Cfg(**{**vars(args), "mode": "compat", "res": "4k", "gops": 8, "unique_gops": 1, "steps": 3, "warmup": 3, "planes": False}),
                  # the headline workload with twice the GOPs per step: the wavefront kernels (I wave) are latency-bound chains whose
                  # duration does not depend on how many pictures run side by side, so a larger batch amortises them
                  "headline_with_32_gops_per_step": Cfg(**{**vars(args), "gops": 32, "steps": 3, "warmup": 3, "planes": False})}
        blocks = {}
        for name, c in others.items():
            try:
                r = measure(c, rank, world, local, dist, torch, extras=False)
                blocks[name] = {"config": config_of(c), "value": r["value"], "unit": "frames/s", "ms_per_step": r["ms_per_step"], "e2e": r["e2e"],
                                "roofline": r["roofline"], "kernel_ms_per_step": r["kernel_ms_per_step"], "content": r["content"]}
            except Exception as e:      # an extra must not lose the headline
                blocks[name] = {"error": "%s: %s" % (type(e).__name__, e)}
        if rank == 0:
            line["configs"] = blocks
    if world > 1 or args.config4:
        try:
            c4 = config4(rank, world, local, dist, torch)
        except Exception as e:
            c4 = {"error": "%s: %s" % (type(e).__name__, e)}
        if rank == 0:
            line["config4"] = c4
    if rank == 0:
        os.sched_setaffinity(0, affinity_before)
        if extras:
            try:
                line["host"] = host_side(args.sample_frames)
            except Exception as e:
                line["host"] = {"error": "%s: %s" % (type(e).__name__, e)}
        if world == 1 and not args.no_cpu_baseline and reference_available():
            try:
                line["cpu_baseline"] = cpu_baseline(args.sample_frames)
            except Exception as e:  # the baseline must not lose the measurement
                line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %s" % e}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
