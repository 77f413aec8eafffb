#!/usr/bin/env python3
"""BASELINE configs[3]: N synthetic events (default 10 M) generated ON THE DEVICE per batch (9.5 TB as binary64 could
not be stored) and pushed through the whole pipeline, sharded over the ranks in contiguous event ranges
(SURVEY.md 8d config 4, 8e): one process per GPU, no collective on the data path -- NCCL only carries the barrier
and the max / sum of the timings and counters.

  python tools/run_cfg4.py [--events 10000000] [--fit-mode 0|1]                                   # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/run_cfg4.py ...

Every rank walks its range in batches of 9 472 events: generate (synth kernel) -> npswf_analyze_batch_device, two
resident buffers alternated so that the generator of batch k+1 runs behind the analysis of batch k on the same
stream.  Timed with CUDA events over the rank's whole range (generation included and reported separately), max over
ranks.  Prints one JSON line (rank 0)."""
import argparse
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
NB, NT, MAXP = 1080, 110, 12


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--events", type=int, default=10_000_000)
    ap.add_argument("--batch", type=int, default=9472)
    ap.add_argument("--config", type=int, default=2, help="pulse content of the events (synth.config_params)")
    ap.add_argument("--fit-mode", type=int, default=0)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import synth
    pkg = importlib.import_module("nps-waveform-analysis_b200")
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    cal = synth.make_calibration()
    h = pkg.NpsWf(cal, devices=[lr], fit_mode=args.fit_mode)
    lo, hi = pkg.shard_range(args.events, rank, world)
    E = args.batch
    p = synth.config_params(args.config)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    st = stream.cuda_stream
    d_spl = torch.from_numpy(h.spline_coeffs()).to(dev); d_tref = torch.from_numpy(cal["timeref"]).to(dev)
    d_kap = torch.from_numpy(cal["kappa"]).to(dev)
    bufs = [(torch.empty((E, NB, NT), dtype=torch.float64, device=dev), torch.empty((E, NB), dtype=torch.int32, device=dev),
             torch.empty((E,), dtype=torch.float64, device=dev)) for _ in range(2)]
    out = dict(wfnpulse=torch.empty((E, NB), dtype=torch.int32, device=dev), wftime=torch.empty((E, NB, MAXP), dtype=torch.float64, device=dev),
               wfampl=torch.empty((E, NB, MAXP), dtype=torch.float64, device=dev), chi2=torch.empty((E, NB), dtype=torch.float64, device=dev),
               timewf=torch.empty((E, NB), dtype=torch.float64, device=dev), amplwf=torch.empty((E, NB), dtype=torch.float64, device=dev),
               status=torch.empty((E, NB), dtype=torch.uint8, device=dev))

    def gen(k, first, n):
        sig, pres, corr = bufs[k % 2]
        synth.generate_device(p, d_spl.data_ptr(), d_tref.data_ptr(), d_kap.data_ptr(), first, n, sig.data_ptr(), 0, pres.data_ptr(),
                              corr.data_ptr(), st)

    def ana(k, n):
        sig, pres, corr = bufs[k % 2]
        h.analyze_device(n, sig.data_ptr(), pres.data_ptr(), corr.data_ptr(), out["wfnpulse"].data_ptr(), out["wftime"].data_ptr(),
                         out["wfampl"].data_ptr(), out["chi2"].data_ptr(), out["timewf"].data_ptr(), out["amplwf"].data_ptr(),
                         out["status"].data_ptr(), stream=st)

    # warm-up (allocations, first launches) on a small batch outside the timed region, and the generator's own rate
    gen(0, lo, min(E, hi - lo)); ana(0, min(E, hi - lo))
    h.sync_device(stream=st)
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record(stream)
    for k in range(4):
        gen(k, lo, min(E, hi - lo))
    g1.record(stream)
    torch.cuda.synchronize()
    gen_ms_per_batch = g0.elapsed_time(g1) / 4
    h.reset_counters()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_wall = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    k = 0
    for first in range(lo, hi, E):
        n = min(E, hi - first)
        gen(k, first, n)
        ana(k, n)
        k += 1
    e1.record(stream)
    torch.cuda.synchronize()
    h.sync_device(stream=st)
    if world > 1:
        dist.barrier()
    wall = time.time() - t_wall
    c = h.counters()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    cnt = torch.tensor([c["n_fit_attempted"], c["n_events"], c["n_fallback"], c["n_fit_ok_retry"], c["n_pulses"]], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    if rank == 0:
        ms = float(t.item())
        fitted, events, fb, retry, pulses = [int(v) for v in cnt.tolist()]
        batches = (hi - lo + E - 1) // E
        print(json.dumps({
            "workload": "BASELINE configs[3]: %d synthetic events (config-%d content), generated on device per batch of %d, "
                        "contiguous event ranges per GPU" % (args.events, args.config, E),
            "n_gpus": world, "fit_mode": {0: "FAST", 1: "MIGRAD", 2: "VM"}.get(args.fit_mode, str(args.fit_mode)), "events": events, "fitted_block_waveforms": fitted,
            "seconds": ms * 1e-3, "wall_seconds": wall, "fitted_block_waveforms_per_s": fitted / (ms * 1e-3),
            "events_per_s": events / (ms * 1e-3), "generator_share": gen_ms_per_batch * batches / ms,
            "fitted_block_waveforms_per_s_excluding_generation": fitted / ((ms - gen_ms_per_batch * batches) * 1e-3),
            "fallback": fb, "retry_ok": retry, "pulses": pulses}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
