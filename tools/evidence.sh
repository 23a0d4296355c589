# usage (on the GPU box, via gpurun): bash tools/evidence.sh <tag>     -> gpurun_out/<tag>/
set -x
TAG=${1:-r02a}
O=gpurun_out/$TAG
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; tail -c 1500 $O/bench.json; tail -5 $O/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; cut -c1-300 $O/bench_ref.json
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs --no-sustained > $O/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs --no-sustained > $O/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_direct_fwd|k_direct_step|k_pv_nodes|k_direct_bwd_poles|k_direct_prep' -s 14 -c 4 -o $O/prof_step -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs --no-sustained > $O/ncu_full.log 2>&1
