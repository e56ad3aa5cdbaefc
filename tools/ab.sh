mkdir -p gpurun_out
run() { env "$@" python tools/ab_frame.py 2>&1 | tail -1; }
MB200_TRACE_MR=440 python -m pytest tests/test_gpu_parity.py tests/test_gpu_render.py -x -q 2>&1 | tail -3
run A=1
run MB200_TRACE_MR=440
run MB200_TRACE_MR=441
run MB200_TRACE_MR=430
run MB200_TRACE_MR=340
run MB200_TRACE_MR=350
run MB200_TRACE_MR=260
run MB200_TRACE_MR=280
