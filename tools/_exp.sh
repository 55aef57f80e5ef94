for d in "" "-DI8_EXP_NOMMA"; do
  touch smrf_b200/csrc/idw_i8.cu; SMRFB_NVCC_DEFS="-DI8_PROFILE $d" python -m smrf_b200.build >/dev/null 2>&1
  echo "== $d"; SMRFB_IDW_KERNEL=i8 IDW_COMPARE=dmma timeout 300 python tools/idw_perf.py 2000 2000 200 720 2>&1 | grep -E "prof|IDW|vs " | tail -3
done
