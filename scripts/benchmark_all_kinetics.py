#!/usr/bin/env python
"""Inference benchmark of the continual models on dummy Kinetics-skeleton data (18 joints, 400 classes), with the
semantics of the reference's ``scripts/benchmark_all_kinetics.py`` (:50-80 there: ``--profile_model
--profile_model_num_runs 10 --forward_mode frame --batch_size 1 --dataset_name dummy_kin``).  Same warm-up / timed-call
protocol as ``benchmark_all_ntu60.py`` in this directory; see there."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import benchmark_all_ntu60 as base  # noqa: E402

base.DEFAULT_DATASET = "dummy_kin"
base.DEFAULT_BATCH, base.DEFAULT_RUNS = 1, 10  # the reference's Kinetics setting (:15,37)

if __name__ == "__main__":
    base.main()
