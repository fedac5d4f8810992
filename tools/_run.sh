for cfg in "12 4" "14 4" "14 6" "13 6" "12 6" "15 4" "12 4"; do
set -- $cfg
POSEFIT_SMALL_WARPS=$1 POSEFIT_DEPTH=$2 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['configs']
print('warps $1 depth $2:', ' '.join('%s %.4f' % (k.split()[0], v['ms']) for k,v in c.items()))"
done
