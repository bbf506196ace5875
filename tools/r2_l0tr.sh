echo YQ_L0_GROUPS=2 YQ_L0_BULK=1; YQ_L0_GROUPS=2 YQ_L0_BULK=1 python tools/probes/l0_trace.py 2>&1 | tail -8
echo YQ_L0_GROUPS=1 YQ_L0_BULK=0; YQ_L0_GROUPS=1 YQ_L0_BULK=0 python tools/probes/l0_trace.py 2>&1 | tail -8
