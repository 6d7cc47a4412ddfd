# ncu --set full captures of the main kernels of the config-2 step (one step's worth of launches each, after the warm-up
# steps), exported as raw CSV; usage: bash tools/ncu_full.sh r2
R=${1:-r2}
CMD="python bench.py --config 2 --steps 1 --warmup 3 --no-graph --no-cpu --no-eager"
cap() {  # name regex skip count
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o /tmp/$1 $CMD > /dev/null 2> gpurun_out/${R}_ncu_$1.err
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/${R}_ncu_full_$1.csv 2>> gpurun_out/${R}_ncu_$1.err
  ls -la /tmp/$1.ncu-rep | tee -a gpurun_out/${R}_ncu_$1.err
}
cap bn_act_bwd '^bn_act_bwd_kernel' 54 18
cap conv_row_wgrad 'conv_row_wgrad_kernel' 48 16
cap conv_row 'conv_row_kernel' 57 19
cap conv_blk 'conv_blk_kernel' 90 30
