mkdir -p gpurun_out
for tool in memcheck synccheck; do   # initcheck takes ~20 min and only reports TMA-stored buffers as uninitialised (profiles/r4e_compute_sanitizer.txt)
  timeout 1500 compute-sanitizer --tool $tool --print-limit 10 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_${tool}_smoke.log 2>&1; echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|smoke ok" gpurun_out/sanitize_${tool}_smoke.log | head -4
  grep -E "Uninitialized|Barrier error|Invalid" -A6 gpurun_out/sanitize_${tool}_smoke.log | head -30
done
