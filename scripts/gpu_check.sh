#!/bin/bash
# gpurun helper: the gpu test suite + a short bench of the strict build; results under gpurun_out/<tag>*
# usage (inside gpurun): bash scripts/gpu_check.sh <tag> [extra bench args]
tag=$1; shift
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_$tag.log
python bench.py --no-cpu-baseline --no-mlv "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
tail -c 1200 gpurun_out/pytest_$tag.log
