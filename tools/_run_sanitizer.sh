# memcheck + racecheck of the sequencer kernel on a small batch (smoke: 8 pairings + 2 Miller loops, all opcodes)
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok|hazard" gpurun_out/sanitizer_$tool.log | head -5
done
