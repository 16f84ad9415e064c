#!/bin/bash
# usage: scratch/quick.sh <tag>   -- GPU tests, short bench, phase breakdown
tag=$1
timeout 150 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 150 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -3 gpurun_out/bench_$tag.err
python -c "
import json; d=json.load(open('gpurun_out/bench_$tag.json')); print(d['value'], d['ms_per_step'], d['roofline']['kernels_ms_per_step'], d['roofline']['whole_step']['frac'], d['decompress']['ms_per_step'], d['parity']['pairs_and_recon_bit_exact'])"
#timeout 100 python scratch/phases.py 32
#timeout 100 python scratch/phases.py 64
