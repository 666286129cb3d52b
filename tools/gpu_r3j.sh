timeout 300 python tests/attn_pair_ab_diag.py old poly0 cur 2>&1 | tail -30
