set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_r1_chrom.json 2> gpurun_out/bench_r1_chrom.err; cut -c1-200 gpurun_out/bench_r1_chrom.json
python bench.py --workload poly > gpurun_out/bench_r1_poly.json 2> gpurun_out/bench_r1_poly.err; cut -c1-200 gpurun_out/bench_r1_poly.json
python bench.py --workload sink > gpurun_out/bench_r1_sink.json 2> gpurun_out/bench_r1_sink.err; cut -c1-200 gpurun_out/bench_r1_sink.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_chrom_v4.csv python bench.py --no-cpu --no-e2e --steps 3 --warmup 3 > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:chrom_kernel -s 3 -c 1 -o gpurun_out/prof_chrom_r1i python bench.py --no-cpu --no-e2e --steps 3 --warmup 3 > gpurun_out/ncu_chromi.log 2>&1
python profiles/ncu_summary.py gpurun_out/prof_chrom_r1i.ncu-rep 25 > gpurun_out/r1_chrom_r1i.ncu_summary.txt 2>&1; head -24 gpurun_out/r1_chrom_r1i.ncu_summary.txt
