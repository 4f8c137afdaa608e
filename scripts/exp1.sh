export CUDA_DEVICE_MAX_CONNECTIONS=32
run() { echo "== $*"; env "$@" timeout 400 python scripts/correct_quick.py F3 $REP 3 2>&1 | grep "^correct\|broker\] tasks" | cut -c1-100 | tail -4; }
export RTK_BROKER_PROFILE=1
REP=32 run A=1
REP=64 run A=1
REP=64 run RTK_CORRECT_INFLIGHT=131072
