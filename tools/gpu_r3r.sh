timeout 600 python tests/model_flake_diag.py 2000 2>&1 | tail -6
timeout 600 python tests/race_stress_diag.py 1500 2>&1 | tail -10
