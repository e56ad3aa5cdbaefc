# ncu_capture.sh NAME COUNT [ENV=VAL ...] -- full capture of COUNT k_trace_sm launches of one bench frame; leaves
# gpurun_out/NAME[_k].md (summary) and NAME[_k]_lines.md (per source line) and deletes the .ncu-rep unless KEEP_REP=1
name=$1; count=$2; shift 2
env "$@" ncu --set full --metrics lts__t_bytes.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum --clock-control none --import-source on -k regex:k_trace_ -c $count -o gpurun_out/$name \
    python tools/prof_frame.py > gpurun_out/${name}_ncu.log 2>&1
for k in $(seq 0 $((count - 1))); do
  python tools/ncu_summary.py gpurun_out/$name.ncu-rep $k > gpurun_out/${name}_$k.md
  python tools/ncu_lines.py gpurun_out/$name.ncu-rep $k 70 > gpurun_out/${name}_${k}_lines.md
done
[ "$KEEP_REP" = 1 ] || rm -f gpurun_out/$name.ncu-rep
