#!/usr/bin/env python
"""Inference benchmark of the continual models on dummy NTU data, with the semantics of the reference's
``scripts/benchmark_all_ntu60.py`` (:52-85 there: ``--profile_model --profile_model_num_runs 100
--forward_mode frame --batch_size 256 --dataset_name dummy_ntu``) and of what ``ride`` does around it:

* ``warm_up`` (models/base.py:144-159): reset the state and push ``receptive_field - padding - 1`` random
  frames through ``forward_step``;
* then time ``num_runs`` calls of ``forward`` on an input of ``stride`` frames (models/base.py:135-142: with
  profiling in frame mode ``input_shape`` becomes ``(C, stride, V, S)``, i.e. one call = one new prediction per
  stream); the state is NOT reset between timed runs (models/base.py:174-175).

Prints, per model, predictions per second (the unit of the paper's Table II), stream-frames per second and the
mean time per call.  GPU only: the hot path has no CPU implementation here.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import continual_skeletons_b200 as cs  # noqa: E402

# the continual models of the reference's loop (:54-61 there) that this package implements; the "*" variants of A-GCN and
# S-TR (coa_gcn_mod, cos_tr_mod) are outside the BASELINE configs
MODELS = {"cost_gcn": cs.CoStGcn, "cost_gcn_mod": cs.CoStGcnMod, "coa_gcn": cs.CoAGcn, "cos_tr": cs.CoSTr}
DEFAULT_DATASET = "dummy_ntu"
DEFAULT_BATCH, DEFAULT_RUNS = 256, 100  # the reference's continual GPU setting (:15-16,52)


def profile_model(name, batch_size, num_runs, dataset_name, device):
    # time_chunk = 1: the protocol is per-step inference; the rings stay at their per-step size (9 / 5 slots)
    model = MODELS[name]({"dataset_name": dataset_name, "forward_mode": "frame", "profile_model": True,
                          "batch_size": batch_size, "time_chunk": 1})
    c, t, v, s = model.input_shape  # t == model.stride in profiling mode
    sample = torch.rand((batch_size, c, t, v, s), device=device)
    model.warm_up(None, sample)
    torch.cuda.synchronize(device)
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = None
    start.record()
    for _ in range(num_runs):
        out = model(sample)
    stop.record()
    torch.cuda.synchronize(device)
    ms = start.elapsed_time(stop) / num_runs
    assert out is not None and tuple(out.shape) == (batch_size, model.num_classes), "a prediction is due on every timed call"
    return {
        "model": name, "batch_size": batch_size, "frames_per_prediction": t, "num_runs": num_runs,
        "ms_per_call": ms, "predictions_per_s": batch_size / (ms * 1e-3), "stream_frames_per_s": batch_size * t / (ms * 1e-3),
        "state_MB_per_stream": model.state_bytes() / batch_size / 1e6,
    }


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--models", nargs="+", default=list(MODELS), choices=list(MODELS))
    ap.add_argument("--batch_size", type=int, default=DEFAULT_BATCH)
    ap.add_argument("--profile_model_num_runs", type=int, default=DEFAULT_RUNS)
    ap.add_argument("--dataset_name", default=DEFAULT_DATASET, choices=["dummy_ntu", "dummy_kin"])
    ap.add_argument("--gpu", type=int, default=0)
    args = ap.parse_args(argv)
    if not torch.cuda.is_available():
        raise SystemExit("the inference benchmark needs a CUDA device")
    device = torch.device("cuda", args.gpu)
    torch.cuda.set_device(device)
    for name in args.models:
        print(json.dumps(profile_model(name, args.batch_size, args.profile_model_num_runs, args.dataset_name, device)), flush=True)


if __name__ == "__main__":
    main()
