# usage: tools/run_variants.sh [--ncu] default name1 name2 ...   (variants from tools/build_variant.sh)
M="gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,launch__shared_mem_config_size,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts.sum"
NCU=0; if [ "$1" = "--ncu" ]; then NCU=1; shift; fi
for v in "$@"; do
  if [ "$v" = default ]; then unset MVG_LIB_PATH; else export MVG_LIB_PATH=mvgformer_b200/variants/libmvg_$v.so; fi
  echo "== $v"; timeout 200 python tools/prof_gather.py
  if [ $NCU = 1 ]; then
  timeout 300 ncu --metrics $M --clock-control none -k regex:"gather_kernel|project_sample_kernel" -s 3 -c 1 python tools/prof_gather.py --iters 2 2>&1 | grep -E "gpu__time|hit_rate|wavefronts|inst_executed|issue_active|lookup_miss|op_ld.sum|config_size"
  fi
done
