#!/bin/bash
# Round 2, job 41: a getter between two replays of the same one-pass graph (e_stale after graph replays).
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_fused.py tests/test_gpu_replay.py -x -q -m gpu -k "getter_between or vacuum_row or replay_from or default_on_large or deferred_stepping_is" ) > gpurun_out/r02_j41.pytest.log 2>&1
tail -n 8 gpurun_out/r02_j41.pytest.log
