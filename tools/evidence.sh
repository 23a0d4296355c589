# usage (on the GPU box, via gpurun): bash tools/evidence.sh <tag> [ncu|bench|all]     -> gpurun_out/<tag>/
#   ncu   : the ncu captures (full-set reports of the direct step, the 1d deck shape and the 2V kernels).  Read prof_step.ncu-rep
#           back with tools/ncu_flops.py -> profiles/ncu_latest.json BEFORE the bench phase, so that the bench line's
#           executed-work counts belong to the binary it times.
#   bench : GPU tests, smoke, bench lines (own arm + reference arm), fit step, 2V timings, ncu launch list
set -x
TAG=${1:-r02h}
PHASE=${2:-all}
O=gpurun_out/$TAG
mkdir -p $O
FLAGS="--steps 2 --warmup 1 --no-cpu-baseline --no-configs --no-sustained"
if [ "$PHASE" = ncu ] || [ "$PHASE" = all ]; then
python bench.py $FLAGS > $O/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_direct_fwd|k_direct_step|k_pv_nodes|k_direct_bwd_poles|k_direct_prep' -s 14 -c 4 -o $O/prof_step -f python bench.py $FLAGS > $O/ncu_full.log 2>&1
python tools/bench_2v.py 64 > $O/plain_2v.log 2>&1 && \
ncu --set full --clock-control none -k regex:'k_ff2v' -s 2 -c 3 -o $O/prof_2v -f python tools/bench_2v.py 64 > $O/ncu_2v.log 2>&1
python tools/bench_1d.py 1024 > $O/plain_1d.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_table_fwd|k_table_bwd' -s 8 -c 3 -o $O/prof_1d -f python tools/bench_1d.py 1024 > $O/ncu_1d.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_1d.csv python tools/bench_1d.py 1024 > $O/ncu_launch_1d.log 2>&1
fi
if [ "$PHASE" = bench ] || [ "$PHASE" = all ]; then
python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
python __graft_entry__.py --smoke > $O/smoke.log 2>&1; tail -1 $O/smoke.log
python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; tail -c 600 $O/bench.json; tail -5 $O/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; cut -c1-200 $O/bench_ref.json
python tools/bench_fit.py 2 > $O/fit_step.txt 2>&1; cat $O/fit_step.txt
python tools/bench_2v.py 1024 > $O/bench_2v.txt 2>&1; cat $O/bench_2v.txt
python bench.py $FLAGS > $O/plain2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py $FLAGS > $O/ncu_launch.log 2>&1
fi
