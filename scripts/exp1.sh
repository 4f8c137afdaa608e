run() { echo "== $*"; env "$@" timeout 300 python scripts/correct_quick.py F3 16 2 2>&1 | grep "^correct" | tail -1; }
export RTK_SPIN_SYNC=1
for ht in 6 8 10 12; do for sv in 2,2,1 3,3,2 4,3,3 4,4,4; do run RTK_HOST_THREADS=$ht RTK_SERVICE_THREADS=$sv; done; done
