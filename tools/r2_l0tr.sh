echo YQ_L0_GROUPS=4; YQ_L0_GROUPS=4 python tools/probes/l0_trace.py 2>&1 | tail -8
echo YQ_L0_GROUPS=2; YQ_L0_GROUPS=2 python tools/probes/l0_trace.py 2>&1 | tail -8
