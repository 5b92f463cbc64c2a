O=gpurun_out/split; mkdir -p $O
cap() { # name split layout target
  DETEXB200_BC6H_SPLIT=$2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:decode_sorted -s 4 -c 1 -f -o $O/ncu_$1 python tools/convert_bench.py --only BPTC_FLOAT --layout $3 --target $4 --steps 3 > $O/ncu_$1.log 2>&1
  python tools/ncu_summary.py $O/ncu_$1.ncu-rep $O/ncu_$1.txt "$1" > /dev/null 2>&1
  ncu -i $O/ncu_$1.ncu-rep --page source --csv > $O/ncu_$1_source.csv 2>/dev/null
  rm -f $O/ncu_$1.ncu-rep
}
cap native_linear_one 0 linear native
cap native_linear_split 1 linear native
cap native_tiled_split 1 tiled native
cap half3_linear_split 1 linear FLOAT_RGB16
ls -la $O
