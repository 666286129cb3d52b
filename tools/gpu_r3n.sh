cp ab/lib_cur.so etude_b200/libetude_b200.so; cp ab/lib_cur_dev.so etude_b200/libetude_b200_dev.so
timeout 600 python tests/race_stress_diag.py 1000 2>&1 | tail -12
