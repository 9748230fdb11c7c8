#!/bin/bash
# development aid: ncu --set full captures of the steady-state wide-path kernels of configs 2 and 5 (run under gpurun)
cap() {  # name, regex, skip, config
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s "$3" -c 1 \
      -o "gpurun_out/r02b_$1" -f python scripts/step_profile.py "$4" 1048576 1 > "gpurun_out/ncu_$1.log" 2>&1
  tail -n 1 "gpurun_out/ncu_$1.log"
}
cap w_rev_cfg2   'w_rev_kernel<\(int\)3, \(int\)1, \(bool\)0>'  6 2
cap w_layer_cfg2 'w_layer_kernel<\(int\)3, \(int\)1, \(int\)0>' 8 2
cap w_layer_cfg5 'w_layer_kernel<\(int\)2, \(int\)0, \(int\)0>' 45 5
cap w_rev_cfg5   'w_rev_kernel<\(int\)2, \(int\)0, \(bool\)0>'  45 5
