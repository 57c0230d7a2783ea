timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for cfg in "1024 1" "640 2" "512 2" "768 1" "416 3" "384 3" "256 3"; do set -- $cfg
  VK_COUNT_THREADS=$1 VK_COUNT_CTAS=$2 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']; print('$cfg', 'count %.1f us'%(k['count']*1e3), 'fold %.1f'%(k['reduce_fold']*1e3), 'total %.1f'%(k['total']*1e3), 'value %.0f'%d['value'])"
done
