mkdir -p gpurun_out
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
run() { # name, env...
  name=$1; shift
  env "$@" timeout 120 python tools/ncu_target.py lin2_ln 3 > gpurun_out/plain_x.log 2>&1 && \
  env "$@" timeout 300 ncu --metrics $M --clock-control none -k regex:gemm2 -s 2 -c 1 --csv --log-file gpurun_out/x_$name.csv python tools/ncu_target.py lin2_ln 3 > gpurun_out/ncu_x.log 2>&1
  echo "== $name rc=$?"; grep -E "dram__bytes|gpu__time|tensor" gpurun_out/x_$name.csv | awk -F'","' '{print $13, $15}'
}
run lsu SAM_GEMM_XLOAD=lsu
run tma SAM_GEMM_XLOAD=tma
run lsucs SAM_GEMM_XLOAD=lsu SAM_GEMM_XCS=1
