for lr in 0.25 0.5 1.0; do for mu in 0.01 0.1 1.0; do
timeout 300 python scripts/c3_solve.py --n 1000000 --maxits 1500 --plain 1 --l-ratio $lr --mu $mu 2>&1 | grep "^{" | python -c "
import sys,json
for ln in sys.stdin:
    o=json.loads(ln); print({k:(round(v,4) if isinstance(v,float) else v) for k,v in o.items() if k in ('l','mu','k_used','k_est','is_nys','setup_s','solve_s','iters','rel_res','plain_iters','plain_rel_res','plain_solve_s')})
"
done; done
