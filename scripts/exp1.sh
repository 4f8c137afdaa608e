export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
run() { echo "== $*"; env "$@" timeout 400 python scripts/correct_quick.py F3 $REP 3 2>&1 | grep "^correct\|broker\]" | cut -c1-200 | tail -6; }
export RTK_BROKER_PROFILE=1
REP=16 run A=1
REP=16 run RTK_SERVICE_THREADS=3,3,2
REP=64 run A=1
