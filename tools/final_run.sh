# regenerates the evidence of a round in one gpurun call: tests, smoke, the four bench lines, the reference arm, launch lists
# of configs 2 and 3, the ncu traffic capture of the dominant entry point.  usage: bash tools/final_run.sh r2
set -x
R=${1:-r2}
(time timeout 600 python -m pytest tests -m gpu -x -q) > gpurun_out/${R}_tests.log 2>&1; tail -3 gpurun_out/${R}_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.log 2>&1; tail -1 gpurun_out/${R}_smoke.log | cut -c1-200
for c in 2 3 4 5; do
  timeout 400 python bench.py --config $c --steps 20 --warmup 3 > gpurun_out/${R}_bench_config$c.json 2> gpurun_out/${R}_bench_config$c.err; tail -1 gpurun_out/${R}_bench_config$c.json | cut -c1-250
done
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference_config2.json 2> gpurun_out/${R}_bench_reference_config2.err; cut -c1-250 gpurun_out/${R}_bench_reference_config2.json
for c in 2 3; do
  timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/${R}_launches_config$c.csv python bench.py --config $c --steps 1 --warmup 1 --no-graph --no-cpu --no-eager > /dev/null 2>&1
done
# DRAM traffic of the dominant entry point of config 2 (b200_bn_act_bwd = reduce + finalize + apply kernels): > one step's launches
timeout 420 ncu --set full --clock-control none --import-source on -k regex:'bn_reduce_kernel|bn_bwd_finalize_kernel|^bn_act_bwd_kernel' -s 124 -c 130 -f -o /tmp/bnbwd python bench.py --config 2 --steps 1 --warmup 3 --no-graph --no-cpu --no-eager > /dev/null 2> gpurun_out/${R}_ncu_bn_bwd.err
ncu -i /tmp/bnbwd.ncu-rep --page raw --csv > gpurun_out/${R}_ncu_full_bn_act_bwd_entry.csv 2>> gpurun_out/${R}_ncu_bn_bwd.err
