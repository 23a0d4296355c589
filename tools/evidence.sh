set -x
mkdir -p gpurun_out/r01j
O=gpurun_out/r01j
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
python bench.py > $O/bench.json 2> $O/bench.err; tail -c 600 $O/bench.json
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; cat $O/bench_ref.json | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_direct_fwd|k_pv_nodes|k_direct_bwd_poles|k_direct_prep' -s 14 -c 4 -o $O/prof_step -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_full.log 2>&1
python tests/parity_report.py > $O/parity.txt 2>&1; tail -20 $O/parity.txt
python tools/bench_configs.py > $O/reference_shapes.txt 2>&1; cat $O/reference_shapes.txt
python tools/bench_fit.py 2 > $O/fit_step.txt 2>&1; tail -2 $O/fit_step.txt
