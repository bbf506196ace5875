export YQ_L0_GROUPS=3
for i in 1 2 3 4 5 6; do ( YQ_NET=tiny YQ_BATCH=64 YQ_WARM=3 YQ_NO_PROFILE_FORWARD=1 timeout 120 python tools/prof_forward.py 2>&1 | grep "Error\|launches per" | head -4 ); done
