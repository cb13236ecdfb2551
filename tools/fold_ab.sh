# A/B of the row count up to which the LN-folded GEMMs finish the LayerNorm statistics themselves (api.cu kFoldStatsMaxRows):
# DAPT masked encoder / full MAE forward (B = 100) and the ViT-B batch sweep at B = 8, 16, 32 (columns: batch, ms, clips/s); one B200.
for t in 8192 32768 65536; do
  echo "FOLD_MAX_ROWS=$t"
  STAD_FOLD_STATS_MAX_ROWS=$t timeout 80 python tools/bench_configs.py --dapt-only --iters 10 2>/dev/null | cut -c1-170
  STAD_FOLD_STATS_MAX_ROWS=$t timeout 80 python -c "
import sys; sys.path.insert(0,'.')
from simple_tad_b200 import efficiency as e
for r in e.batch_sweep('VideoMAE-B', batches=(8,16,32), warmup=5, iters=20, quiet=True): print(r['batch_per_gpu'], round(r['ms'],3), round(r['clips_per_s'],1))
" 2>/dev/null
done
