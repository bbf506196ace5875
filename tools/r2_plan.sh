YQ_DEBUG_PLAN=1 YQ_NET=yolov3 YQ_BATCH=2 YQ_WARM=0 YQ_NO_PROFILE_FORWARD=1 python tools/prof_forward.py 2>&1 | grep "yq plan" | head -60
