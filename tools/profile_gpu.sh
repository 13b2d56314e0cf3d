#!/bin/bash
# Run on the GPU box (gpurun): ncu launch list of one bench run + one `--set full` capture of the kernels that carry
# the configuration's cycle.  usage: tools/profile_gpu.sh <config> <tag>
cfg=${1:-3}; tag=${2:-r02}
mkdir -p gpurun_out
case $cfg in
  3) rx='k_cycle_flat|k_cycle_root'; skip=6; cnt=2;;
  2) rx='k_lone|k_nominate|k_scan_roots|k_scatter|k_rank|k_admit_lone'; skip=24; cnt=6;;
  4) rx='k_rank_keys|k_columns|k_frl_fill|k_root_recs|k_cells_mark|k_search_cells_grouped|k_nominate_walk|k_admit|k_tree|k_rank'; skip=60; cnt=16;;
  5) rx='k_tas_leaf|k_tas_reduce|k_tas_select'; skip=12; cnt=4;;
esac
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv \
    --log-file gpurun_out/${tag}_launches_cfg${cfg}.csv python bench.py --config $cfg --steps 2 --warmup 3 --no-cpu-baseline --no-drain > gpurun_out/${tag}_launches_cfg${cfg}.log 2>&1
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k "regex:^(void )?($rx)" -s $skip -c $cnt \
    -o gpurun_out/${tag}_full_cfg${cfg} -f python bench.py --config $cfg --steps 2 --warmup 3 --no-cpu-baseline --no-drain > gpurun_out/${tag}_full_cfg${cfg}.log 2>&1
# summarise on the box (the reports carry the source pages and are too large to bring back) and keep only the text
python tools/summarize_ncu.py gpurun_out/${tag}_full_cfg${cfg}.ncu-rep cfg${cfg} ${tag} > gpurun_out/${tag}_summ_cfg${cfg}.log 2>&1
cp profiles/${tag}_ncu_cfg${cfg}.txt gpurun_out/ 2>/dev/null
cp profiles/traffic.json gpurun_out/${tag}_traffic_after_cfg${cfg}.json 2>/dev/null
[ "${KEEP_REP:-0}" = 1 ] || rm -f gpurun_out/${tag}_full_cfg${cfg}.ncu-rep
ls -la gpurun_out/ | grep ${tag}_ | tail -8
