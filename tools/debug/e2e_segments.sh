for cfg in "8 67108864" "16 67108864" "32 33554432" "24 67108864"; do set -- $cfg; echo "max_segments=$1 seg_bytes=$2"; COVT_MAX_SEGMENTS=$1 COVT_SEG_BYTES=$2 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e-host 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('  step %.2f ms  e2e %.2f ms  %.2f GB/s  frac_of_ceiling %.3f  h2d_copy %.2f ms' % (d['ms_per_step'], e['ms_per_step'], e['value'], e['fraction_of_h2d_ceiling'], e['h2d_copy_ms_per_step']))"; done
