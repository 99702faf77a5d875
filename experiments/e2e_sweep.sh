# end-to-end sub-batch schedules (cumulative fractions): prints device ms/step, pipelined e2e ms/step, one-step-at-a-time e2e ms/step
for ch in "1" "0.5,1" "0.25,0.5,0.75,1" "0.125,0.375,0.625,0.875,1" "0.03125,0.125,0.25,0.4,0.55,0.7,0.85,1"; do
  SRB200_E2E_CHUNKS=$ch timeout 200 python bench.py --no-cpu-baseline --no-subrecords 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$ch', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), round(d['e2e']['one_step_at_a_time']['ms_per_step'],3))"
done
