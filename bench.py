#!/usr/bin/env python3
"""bench.py — fitted block-waveforms/s of the NPS waveform hot path on N B200s (one process per GPU).

Workload (BASELINE.json configs[1]): synthetic events, 1-3 pulses/block with pile-up, all 1080 blocks
present, generated ON DEVICE (synth/) into resident HBM batches.  A "step" is one pass of the full
pipeline (matched filter -> TSpectrum search -> 3x3 cluster threshold -> template fit) over one batch.

  value : whole-job fitted block-waveforms/s with inputs resident in HBM (device-timed, max over ranks)
  e2e   : the same metric through the reference-facing C-ABI call npswf_analyze_batch with pinned HOST
          buffers; H2D of the inputs and D2H of every output are inside the timed region
  roofline / stages : per-stage CUDA-event times (library profiling hooks record events on the launching
          stream) from a second pass over the same batches inside this script: with the hooks on, the library
          serialises its streams so that every stage is timed alone; the timed region of `value` runs with
          the hooks off (fit kernels on side streams, chunks overlapped).  Algorithmic bytes per SURVEY.md §8(d)
  cpu_baseline : the CPU oracle (a port/restatement of the reference path — ROOT is not installable)
          on the box's host cores, on a bounded sample of the same workload

--impl reference times that CPU restatement alone (all host threads) and prints the same JSON line.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NB, NT, MAXP = 1080, 110, 12
METRIC = "fitted block-waveforms/sec"
UNIT = "block-waveforms/s"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload_string(batch, steps):
    return ("BASELINE configs[1]: synthetic events x 1080 blocks x 110 samples, 1-3 pulses/block with pile-up "
            "(A~logU[3,500] mV, sep>=3 bins, sigma=0.30 mV, 12-bit ADC lattice), all blocks present; "
            "%d events/step/GPU, %d steps" % (batch, steps))


def run_reference(args, rank, world):
    """CPU restatement of the reference path (oracle port), all host threads, bounded sample per step."""
    if rank != 0:
        return
    import oracle
    import synth
    cal = synth.make_calibration()
    orc = oracle.Oracle(cal)
    spl = orc.spline_coeffs()
    threads = os.cpu_count() or 1
    n_ev = args.ref_events if args.ref_events > 0 else max(96, 6 * threads)   # >= 6 events per thread: the pool stays loaded
    p = synth.config_params(2)
    times, fitted = [], 0
    for s in range(args.warmup + args.steps):
        ev = synth.generate_host(p, spl, cal, 10_000_000 + s * n_ev, n_ev, n_threads=threads)
        t0 = time.perf_counter()
        r = orc.analyze_batch(ev["signal"], ev["pres"], ev["corr_time_HMS"], n_threads=threads)
        dt = time.perf_counter() - t0
        if s >= args.warmup:
            times.append(dt)
            fitted += int(((r["status"] & 28) > 0).sum())
    total = float(sum(times))
    value = fitted / total
    # the cost-faithful mode (spline rebuilt per fit, one mutex around the peak search) on one more sample
    orc_f = oracle.Oracle(cal, flags=oracle.FLAG_FAITHFUL_COST)
    ev = synth.generate_host(p, spl, cal, 10_000_000, n_ev, n_threads=threads)
    t0 = time.perf_counter()
    rf = orc_f.analyze_batch(ev["signal"], ev["pres"], ev["corr_time_HMS"], n_threads=threads)
    faithful = float(((rf["status"] & 28) > 0).sum()) / (time.perf_counter() - t0)
    # BASELINE configs[0]: single pulse per block with EnableImplicitMT(4) (T2:313, README:33): 4 threads, bounded sample
    ev0 = synth.generate_host(synth.config_params(1), spl, cal, 20_000_000, 48, n_threads=threads)
    t0 = time.perf_counter()
    r0 = orc.analyze_batch(ev0["signal"], ev0["pres"], ev0["corr_time_HMS"], n_threads=4)
    cfg0 = float(((r0["status"] & 28) > 0).sum()) / (time.perf_counter() - t0)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, len(times)), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_string(n_ev, args.steps),
                   "note": "CPU restatement of npsWF.C's path (TSpectrum + Minuit2-Migrad restated; ROOT is not "
                           "installable here); std::thread pool over events mirrors EnableImplicitMT"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d events/step x %d steps of the same workload" % (n_ev, args.steps),
                         "faithful_cost_mode": {"value": faithful, "unit": UNIT,
                                                "sample": "%d events, spline rebuilt per fit + search mutex" % n_ev},
                         "config0_4_threads": {"value": cfg0, "unit": UNIT, "cores": 4,
                                               "sample": "BASELINE configs[0]: 48 of its 1 000 single-pulse events, 4 threads (EnableImplicitMT(4))"}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch-events", type=int, default=9472,
                    help="events per step per GPU: one npswf_analyze_batch_device call, which the library cuts into two chunks of 4 736")
    ap.add_argument("--e2e-events", type=int, default=4736, help="events per end-to-end step per GPU")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--migrad-steps", type=int, default=2, help="steps of the MIGRAD-fit-mode leg (0 = skip)")
    ap.add_argument("--vm-steps", type=int, default=4, help="steps of the VM-fit-mode leg (0 = skip)")
    ap.add_argument("--stage-steps", type=int, default=4, help="steps of the serialised stage-profiling pass")
    ap.add_argument("--ref-events", type=int, default=0, help="--impl reference: events per step (0 = 6 per host thread, at least 96)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import synth
    pkg = importlib.import_module("nps-waveform-analysis_b200")

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    cal = synth.make_calibration()
    h = pkg.NpsWf(cal, devices=[local_rank])
    E = args.batch_events
    p = synth.config_params(2)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    st = stream.cuda_stream
    d_spl = torch.from_numpy(h.spline_coeffs()).to(dev)
    d_tref = torch.from_numpy(cal["timeref"]).to(dev)
    d_kap = torch.from_numpy(cal["kappa"]).to(dev)

    n_buf = 2   # two distinct resident batches, alternated; each is far larger than the 126 MB L2
    bufs = []
    for b in range(n_buf):
        sig = torch.empty((E, NB, NT), dtype=torch.float64, device=dev)
        pres = torch.empty((E, NB), dtype=torch.int32, device=dev)
        corr = torch.empty((E,), dtype=torch.float64, device=dev)
        ev0 = (rank * n_buf + b) * E     # contiguous, disjoint event ranges per rank (weak scaling)
        synth.generate_device(p, d_spl.data_ptr(), d_tref.data_ptr(), d_kap.data_ptr(), ev0, E, sig.data_ptr(), 0,
                              pres.data_ptr(), corr.data_ptr(), st)
        bufs.append((sig, pres, corr))
    out = dict(wfnpulse=torch.empty((E, NB), dtype=torch.int32, device=dev),
               wftime=torch.empty((E, NB, MAXP), dtype=torch.float64, device=dev),
               wfampl=torch.empty((E, NB, MAXP), dtype=torch.float64, device=dev),
               chi2=torch.empty((E, NB), dtype=torch.float64, device=dev),
               timewf=torch.empty((E, NB), dtype=torch.float64, device=dev),
               amplwf=torch.empty((E, NB), dtype=torch.float64, device=dev),
               status=torch.empty((E, NB), dtype=torch.uint8, device=dev))
    torch.cuda.synchronize()

    def step(i):
        sig, pres, corr = bufs[i % n_buf]
        h.analyze_device(E, sig.data_ptr(), pres.data_ptr(), corr.data_ptr(), out["wfnpulse"].data_ptr(),
                         out["wftime"].data_ptr(), out["wfampl"].data_ptr(), out["chi2"].data_ptr(),
                         out["timewf"].data_ptr(), out["amplwf"].data_ptr(), out["status"].data_ptr(), stream=st)

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    h.sync_device(stream=st)
    h.reset_counters()
    h.set_profiling(False)
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        step(args.warmup + i)
    e1.record(stream)
    torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop()
    h.sync_device(stream=st)
    ms_total = e0.elapsed_time(e1)
    ctr = h.counters()
    fitted_local = ctr["n_fit_attempted"]
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    cnt = torch.tensor([fitted_local, ctr["n_block_waveforms"], ctr["n_pulses"], ctr["n_fit_iterations"],
                        ctr["n_fallback"], ctr["n_fit_ok_retry"]], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms_total = float(t.item())
    fitted, blocks, pulses, iters, n_fb, n_retry = [int(v) for v in cnt.tolist()]
    value = fitted / (ms_total * 1e-3)

    # ---- stage pass: same batches, profiling hooks on (streams serialised, every stage timed alone)
    h.reset_counters()
    h.stage_times(reset=True)
    h.set_profiling(True)
    for i in range(args.stage_steps):
        step(i)
    h.sync_device(stream=st)
    h.set_profiling(False)
    stages = h.stage_times(reset=True)
    sctr = h.counters()
    fp64_peak = h.fp64_peak_gflops() if rank == 0 else 0.0

    # ---- end to end through npswf_analyze_batch: pinned host buffers, H2D + D2H inside the timed region
    e2e = None
    migrad = None
    vm = None

    def e2e_leg(hh, call, steps):
        """One warm-up call, then `steps` synchronous calls timed on the host clock between barriers; max over ranks."""
        call()
        hh.reset_counters()
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            call()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        c = hh.counters()
        te = torch.tensor([dt], dtype=torch.float64, device=dev)
        ce = torch.tensor([c["n_fit_attempted"]], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
            dist.all_reduce(ce, op=dist.ReduceOp.SUM)
        return float(ce.item()) / float(te.item()), 1e3 * float(te.item()) / steps

    if not args.no_e2e:
        Ee = args.e2e_events      # the same step at every N (weak scaling): ~6.8 GB of pinned host memory per rank
        hs = pkg.pinned_empty((Ee, NB, NT), np.float64)
        hp = pkg.pinned_empty((Ee, NB), np.int32)
        hc = pkg.pinned_empty((Ee,), np.float64)
        for o0 in range(0, Ee, E):          # the same synthetic events as the resident batches
            b = bufs[(o0 // E) % n_buf]
            n = min(E, Ee - o0)
            hs[o0:o0 + n] = b[0][:n].cpu().numpy()
            hp[o0:o0 + n] = b[1][:n].cpu().numpy()
            hc[o0:o0 + n] = b[2][:n].cpu().numpy()
        ho = h.alloc_outputs(Ee, pinned=True)
        host_in = hs.nbytes + hp.nbytes + hc.nbytes
        d2h = sum(v.nbytes for v in ho.values())
        # what the host link gives every rank when all ranks upload at once: 1 GiB of the pinned trace buffer, raw
        probe_t = torch.empty((1 << 27,), dtype=torch.float64, device=dev)
        src = torch.from_numpy(np.asarray(hs).reshape(-1)[:1 << 27])
        probe_t.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        barrier()
        pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pe0.record(stream)
        for _ in range(3):
            probe_t.copy_(src, non_blocking=True)
        pe1.record(stream)
        torch.cuda.synchronize()
        h2d_gbs = torch.tensor([3 * (1 << 30) / (pe0.elapsed_time(pe1) * 1e-3) / 1e9], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(h2d_gbs, op=dist.ReduceOp.MIN)
        del probe_t, src
        h.analyze(hs, hp, hc, out=ho)          # allocates the staging buffers
        ps0 = h.host_packing_stats()
        v, ms = e2e_leg(h, lambda: h.analyze(hs, hp, hc, out=ho), args.e2e_steps)
        ps1 = h.host_packing_stats()
        packed_in = (ps1["packed_input_bytes"] - ps0["packed_input_bytes"]) // (args.e2e_steps + 1)
        h2d = host_in - (packed_in * 3) // 4        # what crossed PCIe: the packed part of the traces is a quarter of its size
        up_rate, bound_cpus = h.host_upload_rate()
        e2e = {"value": v, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "events_per_step_per_gpu": Ee, "steps": args.e2e_steps,
               "host_input_bytes_per_step": int(host_in),
               "input": "f64 [E][1080][110] (the reference's Double_t layout), pinned host memory",
               "transport": "library default: host threads rewrite lattice traces as int16 counts when that reproduces every "
                            "double (lossless, checked per sample), raw doubles otherwise; %.0f %% of the trace bytes packed at "
                            "%.1f GB/s" % (100.0 * packed_in / max(1, hs.nbytes), ps1["pack_gb_per_s"]),
               "ms_per_step": ms,
               "host_link": {"h2d_gbs_per_rank_all_ranks_uploading": float(h2d_gbs.item()),
                             "note": "plain pinned H2D copy of 1 GiB, all ranks at once, slowest rank",
                             "library_measured_raw_upload_gbs": up_rate, "numa_bound_cpus": bound_cpus}}
        # the same call with the packing off: every trace crosses PCIe as binary64
        h.set_host_packing(0)
        v, ms = e2e_leg(h, lambda: h.analyze(hs, hp, hc, out=ho), max(2, args.e2e_steps // 3))
        h.set_host_packing(1)
        e2e["f64_raw_transport"] = {"value": v, "unit": UNIT, "h2d_bytes_per_step": int(host_in), "d2h_bytes_per_step": int(d2h),
                                    "ms_per_step": ms, "input": "same call, npswf_set_host_packing(0): PCIe-bound"}
        # the same analysis through npswf_analyze_batch_flat: wfampl / wftime come back in the reference's truncated
        # layout (T2:1289-1296), packed on the device -- a fraction of the D2H bytes, no flatten pass on the host
        hf = h.alloc_flat_outputs(Ee, Ee * NB * 4, pinned=True)
        v, ms = e2e_leg(h, lambda: h.analyze_flat(hs, hp, hc, out=hf), max(2, args.e2e_steps // 3))
        fixed = sum(hf[k].nbytes for k in ("wfnpulse", "chi2", "timewf", "amplwf", "status"))
        d2h_flat = int(fixed + 16 * hf["n_pulses"] + 12 * Ee)
        e2e["flat_outputs"] = {"value": v, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": d2h_flat,
                               "ms_per_step": ms,
                               "input": "same f64 host input; npswf_analyze_batch_flat (pulses packed on the device)"}
        # the int16 ADC-count ABI (exact on the 1000/4096 mV lattice): 4x less PCIe traffic in
        hk = pkg.pinned_empty((Ee, NB, NT), np.int16)
        hk[...] = np.rint(hs / synth.LSB).astype(np.int16)
        v, ms = e2e_leg(h, lambda: h.analyze_i16(hk, synth.LSB, hp, hc, out=ho), max(2, args.e2e_steps // 3))
        h2d_i16 = int(hk.nbytes + hp.nbytes + hc.nbytes)
        e2e["i16_abi"] = {"value": v, "unit": UNIT, "h2d_bytes_per_step": h2d_i16, "d2h_bytes_per_step": int(d2h), "ms_per_step": ms,
                          "input": "int16 ADC counts [E][1080][110] + lsb (npswf_analyze_batch_i16), pinned host memory"}
        # the multi-GPU transport: int16 counts in, the reference's packing out (npswf_analyze_batch_flat_i16)
        v, ms = e2e_leg(h, lambda: h.analyze_flat_i16(hk, synth.LSB, hp, hc, out=hf), args.e2e_steps)
        e2e["i16_flat"] = {"value": v, "unit": UNIT, "h2d_bytes_per_step": h2d_i16, "d2h_bytes_per_step": d2h_flat, "ms_per_step": ms,
                           "bytes_per_block_waveform": (h2d_i16 + d2h_flat) / float(Ee * NB),
                           "input": "npswf_analyze_batch_flat_i16: int16 counts in, pulses packed on the device out"}

    # ---- one process, all GPUs: the handle's own event sharding (n_devices = N, one host thread per device inside the
    # call) instead of one process per GPU, same transport as i16_flat.  Rank 0 runs it while the other ranks wait.
    if not args.no_e2e and world > 1:
        barrier()
        if rank == 0:
            hN = pkg.NpsWf(cal, devices=list(range(world)))
            EN = Ee * world
            hkN = pkg.pinned_empty((EN, NB, NT), np.int16)
            hpN = pkg.pinned_empty((EN, NB), np.int32)
            hcN = pkg.pinned_empty((EN,), np.float64)
            for r in range(world):
                hkN[r * Ee:(r + 1) * Ee] = hk
                hpN[r * Ee:(r + 1) * Ee] = hp
                hcN[r * Ee:(r + 1) * Ee] = hc
            hfN = hN.alloc_flat_outputs(EN, EN * NB * 4, pinned=True)
            hN.analyze_flat_i16(hkN, synth.LSB, hpN, hcN, out=hfN)
            hN.reset_counters()
            t0 = time.perf_counter()
            for _ in range(3):
                hN.analyze_flat_i16(hkN, synth.LSB, hpN, hcN, out=hfN)
            dtN = time.perf_counter() - t0
            e2e["in_process_all_gpus"] = {"value": hN.counters()["n_fit_attempted"] / dtN, "unit": UNIT, "ms_per_step": 1e3 * dtN / 3,
                                          "events_per_step": EN,
                                          "input": "ONE process, handle with n_devices = %d (contiguous event ranges, one host thread per "
                                                   "device inside npswf_analyze_batch_flat_i16), int16 in / flat out" % world}
            del hN, hkN, hpN, hcN, hfN
        barrier()

    # ---- the other two fit modes on the same resident batches and through the same host call, fewer steps:
    # MIGRAD = the reference's own minimiser on the device (outputs bit-identical to the oracle, a step takes ~5x longer);
    # VM = Migrad's recursion with analytic derivatives (agrees with Migrad on 99.99 % of 1-3-pulse fits)
    def mode_leg(mode, steps, note):
        hm = pkg.NpsWf(cal, devices=[local_rank], fit_mode=mode)

        def mstep(i):
            sig, pres, corr = bufs[i % n_buf]
            hm.analyze_device(E, sig.data_ptr(), pres.data_ptr(), corr.data_ptr(), out["wfnpulse"].data_ptr(),
                              out["wftime"].data_ptr(), out["wfampl"].data_ptr(), out["chi2"].data_ptr(),
                              out["timewf"].data_ptr(), out["amplwf"].data_ptr(), out["status"].data_ptr(), stream=st)
        mstep(0)
        hm.sync_device(stream=st)
        hm.reset_counters()
        if mode == pkg.FIT_VM:
            hm.vm_reasons()                                    # clears the hand-off tallies of the warm-up step
        barrier()
        torch.cuda.synchronize()
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m0.record(stream)
        for i in range(steps):
            mstep(1 + i)
        m1.record(stream)
        torch.cuda.synchronize()
        hm.sync_device(stream=st)
        mc = hm.counters()
        tm = torch.tensor([m0.elapsed_time(m1)], dtype=torch.float64, device=dev)
        cm = torch.tensor([mc["n_fit_attempted"], mc["n_fit_evals"], mc["n_fit_ok_retry"], mc["n_fallback"]], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            dist.all_reduce(cm, op=dist.ReduceOp.SUM)
        leg = {"value": float(cm[0].item()) / (float(tm.item()) * 1e-3), "unit": UNIT, "steps": steps,
               "ms_per_step": float(tm.item()) / steps,
               "chi2_evaluations_per_fit": float(cm[1].item()) / max(1.0, float(cm[0].item())),
               "retry_ok": int(cm[2].item()), "fallback": int(cm[3].item()), "note": note}
        if mode == pkg.FIT_VM:
            r = hm.vm_reasons()
            leg["handed_to_exact_kernels"] = {"evaluation_limit": int(r[0]), "edm_above_tolerance": int(r[3]), "other": int(r[1] + r[2] + r[4])}
        if not args.no_e2e:
            v, ms = e2e_leg(hm, lambda: hm.analyze(hs, hp, hc, out=ho), 2)
            leg["e2e"] = {"value": v, "unit": UNIT, "ms_per_step": ms, "h2d_bytes_per_step": e2e["h2d_bytes_per_step"],
                          "d2h_bytes_per_step": e2e["d2h_bytes_per_step"], "input": "same f64 host call as e2e"}
        del hm
        return leg

    if args.migrad_steps > 0:
        migrad = mode_leg(pkg.FIT_MIGRAD, args.migrad_steps,
                          "fit_mode = NPSWF_FIT_MIGRAD: Minuit2-Migrad re-implemented on the device (numerical gradients, strategy "
                          "1 -> 2), every output bit-identical to the CPU oracle (tests/test_gpu_migrad.py); same resident batches")
    if args.vm_steps > 0:
        vm = mode_leg(pkg.FIT_VM, args.vm_steps,
                      "fit_mode = NPSWF_FIT_VM: Migrad's line search / Davidon update / EDM stop with analytic derivatives, one "
                      "thread per fit; fits leaving the common path go to the exact Migrad kernels; within tolerance of the "
                      "oracle's Migrad on 99.99 % (1-3 pulses) / 99.5 % (<= 12 pulses) of fits (tests/test_gpu_migrad.py)")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel + per-stage table (rank 0's CUDA-event times inside the timed region)
    peaks, peak_src = _peaks()
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    chunks = max(1, stages["chunks"])
    units_local = sctr["n_block_waveforms"]
    nbar = sctr["n_pulses"] / max(1, sctr["n_present"])
    # algorithmic bytes per block-waveform, SURVEY.md §8(d) (FP64 input ABI, s = 8)
    alg = {
        "front": 880.0 + 1.0,                  # matched filter + 3x3 threshold: each sample once, 1 flag byte
        "search": 440.0 + 4.0 + 16.0 * nbar,   # reads the float MF spectrum, writes wfnpulse + (t, A) per pulse
        "fit": 880.0 + 16.0 * nbar + 13.0 + 16.0 * nbar,  # per FITTED block: trace + seeds in, chi2/status/(t, A) out
    }
    stage_units = {"front": units_local, "search": units_local, "fit": sctr["n_fit_attempted"]}
    stage_rows = {}
    for k in ("front", "search", "fit"):
        ms = stages[k + "_ms"]
        gbs = stage_units[k] * alg[k] / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        stage_rows[k] = {"ms_per_step": ms / args.stage_steps, "share": ms / max(1e-9, sum(stages[s + "_ms"] for s in ("front", "search", "fit"))),
                         "alg_bytes_per_unit": alg[k], "achieved_gbs": gbs, "frac_of_hbm": gbs / hbm,
                         "units_per_s": stage_units[k] / (ms * 1e-3) if ms > 0 else 0.0}
    dom = max(("front", "search", "fit"), key=lambda k: stages[k + "_ms"])
    launches_per_chunk = {"front": 1, "search": 1, "fit": 14}
    # dram__bytes_read.sum + dram__bytes_write.sum per block-waveform, from the `ncu --set full` capture of this round's
    # final build (profiles/r2c_ncu_full_search_front_fit.csv: launches of 1 184 events = 1 278 720 block-waveforms)
    ncu_traffic_per_unit = {"front": (1.132395e9 + 539.443e6) / 1278720.0, "search": (0.914046e9 + 280.581e6) / 1278720.0}
    units_per_launch = units_local / chunks
    traffic = ncu_traffic_per_unit[dom] * units_per_launch if dom in ncu_traffic_per_unit else None
    roofline = {"kernel": {"front": "front_kernel", "search": "search_kernel", "fit": "fit_thread_kernel<1,2> + fit_small_kernel + fit_kernel<25> (14 launches)"}[dom],
                "bound": "hbm", "achieved": stage_rows[dom]["achieved_gbs"], "peak": hbm, "unit": "GB/s",
                "frac": stage_rows[dom]["achieved_gbs"] / hbm, "traffic": traffic,
                "traffic_source": "ncu --set full, profiles/r2c_ncu_full_search_front_fit.csv (bytes per block-waveform x block-waveforms per launch)",
                "peak_source": peak_src,
                "note": "dominant stage is FP64-pipe bound, not HBM bound (bit-faithful FP64 TSpectrum / FP64 LM); "
                        "see stages, fp64_peak_gflops_measured and DESIGN.md",
                "measured_in": "serialised stage pass of %d steps inside bench.py (CUDA events on the launching stream)" % args.stage_steps,
                "avg_launch_ms": stages[dom + "_ms"] / chunks / launches_per_chunk[dom]}
    # The search kernel is bound by the FP64 pipe, not by HBM: the same launch expressed in algorithmic FP64
    # operations (39 000 add/mul/fma/compare per spectrum, an FMA counted once) against the
    # instruction-rate peak of the pipe (measured FMA-chain GFLOP/s / 2).
    if dom == "search" and fp64_peak > 0:
        ops_unit = 39000.0
        ach = stage_rows["search"]["units_per_s"] * ops_unit / 1e12
        roofline["fp64"] = {"ops_per_unit": ops_unit, "achieved_tops": ach, "peak_tops": fp64_peak / 2e3,
                            "frac": ach / (fp64_peak / 2e3), "unit": "T FP64 instr-lanes/s",
                            "note": "ops_per_unit = FP64 operations of the reference's arithmetic (SearchHighRes as ROOT runs it); the "
                                    "kernel's fused pass executes 893 FP64 warp-instructions (24 500 lane-operations) per spectrum for "
                                    "them (ncu, profiles/r2d_source_lines_search.txt: FP64 pipe 41.8 %, issue slots 65.9 %) and repeats "
                                    "0.02 % of the spectra with the exact arithmetic"}
    # fit-stage arithmetic: SURVEY §8(d) flops_iter(N) with the mean multiplicity
    n_mean = pulses / max(1, fitted)
    Pm = 1 + 2 * n_mean
    flops_iter = 90 * (16 * n_mean + Pm * (Pm + 1) + 2 * Pm + 4) + Pm ** 3 / 3 + 2 * Pm ** 2
    fit_gflops = (sctr["n_fit_iterations"] * flops_iter) / (stages["fit_ms"] * 1e-3) / 1e9 if stages["fit_ms"] > 0 else 0.0

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        import oracle
        orc = oracle.Oracle(cal)
        threads = os.cpu_count() or 1
        probe = max(2, threads)   # one event per thread: the probe sees the parallel rate
        cap = min(1184, E)
        sig = bufs[0][0][:cap].cpu().numpy(); prs = bufs[0][1][:cap].cpu().numpy(); cor = bufs[0][2][:cap].cpu().numpy()
        t0 = time.perf_counter()
        orc.analyze_batch(sig[:probe], prs[:probe], cor[:probe], n_threads=threads)
        per_ev = (time.perf_counter() - t0) / probe
        n_s = int(max(threads, min(cap, args.cpu_seconds / max(per_ev, 1e-6))))
        n_s = min(cap, max(probe, n_s))
        t0 = time.perf_counter()
        r = orc.analyze_batch(sig[:n_s], prs[:n_s], cor[:n_s], n_threads=threads)
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": float(((r["status"] & 28) > 0).sum()) / dt, "unit": UNIT, "cores": threads,
                        "kind": "port", "seconds": dt,
                        "sample": "first %d events of the resident batch (same workload), oracle with %d threads" % (n_s, threads)}
        # the cost-faithful mode (spline rebuilt per fit, one mutex around the peak search), smaller sample
        n_f = max(probe, n_s // 4)
        orc_f = oracle.Oracle(cal, flags=oracle.FLAG_FAITHFUL_COST)
        t0 = time.perf_counter()
        rf = orc_f.analyze_batch(sig[:n_f], prs[:n_f], cor[:n_f], n_threads=threads)
        dtf = time.perf_counter() - t0
        cpu_baseline["faithful_cost_mode"] = {"value": float(((rf["status"] & 28) > 0).sum()) / dtf, "unit": UNIT,
                                              "seconds": dtf,
                                              "sample": "first %d events, spline rebuilt per fit + search mutex" % n_f}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic (generated on device, Philox-keyed by seed/event/block)",
        "config": {"workload": workload_string(E, args.steps), "batch_events_per_gpu": E,
                   "events_total": E * args.steps * world,
                   "block_waveforms_per_s": blocks / (ms_total * 1e-3),
                   "fitted_fraction": fitted / max(1, blocks), "mean_pulses_per_fit": n_mean,
                   "fit_iterations_mean": iters / max(1, fitted), "fallback": n_fb, "retry_ok": n_retry,
                   "l2": "inputs larger than L2: %.2f GB of traces per step, two resident batches alternated" % (E * NB * NT * 8 / 1e9)},
        "clocks": clocks, "e2e": e2e, "fit_mode": "NPSWF_FIT_FAST (Levenberg-Marquardt, library default): 99.86 % of this workload's fits within tolerance of the oracle's Migrad; fit_mode_vm: 99.989 %; fit_mode_migrad: bit-identical (tests/test_gpu_migrad.py)", "fit_mode_migrad": migrad, "fit_mode_vm": vm,
        # front, search, compact and 18 fit kernels per chunk; chunks per step as counted by the library in the stage pass
        "gpu_launches": int(args.steps * (chunks // max(1, args.stage_steps)) * 21),
        "roofline": roofline, "stages": stage_rows,
        "fit_fp64": {"gflops": fit_gflops, "flops_per_iter_model": flops_iter, "fp64_peak_gflops_measured": fp64_peak,
                     "frac_of_fp64_peak": fit_gflops / fp64_peak if fp64_peak > 0 else None,
                     "note": "SURVEY 8(d) flop model x measured accepted iterations / fit-stage time; peak = FMA-chain microbenchmark on this GPU"},
        "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
