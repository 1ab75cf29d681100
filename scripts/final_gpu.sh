python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_final2.log
python bench.py > gpurun_out/bench_full_final2.json 2> gpurun_out/bench_full_final2.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-mlv > gpurun_out/r02_ncu_bench.log 2>&1
ncu --set full --clock-control none -k regex:k_ -s 171 -c 57 -o /tmp/r02.ncu-rep python scripts/perf_dump.py 9504 6336 0.4 > gpurun_out/r02_perf_dump.txt 2>&1
python scripts/ncu_summary.py /tmp/r02.ncu-rep > gpurun_out/r02_all_summary.txt 2> gpurun_out/r02_all_summary.err
cat gpurun_out/pytest_final2.log; tail -c 400 gpurun_out/bench_full_final2.json; wc -l gpurun_out/r02_launches.csv gpurun_out/r02_all_summary.txt
