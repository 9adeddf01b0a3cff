for shape in "256 128 128 3" "128 128 128 3"; do set -- $shape; cin=$1; cout=$2; res=$3; k=$4
for dm in 0 16 17; do echo "== $shape DBGMODE=$dm"; PASTA_B200_CONV_DBGMODE=$dm timeout 60 python tools/conv_timeline.py --cin $cin --cout $cout --res $res --k $k | grep -E "main loop|total|waiting on A|kernel span"; done; done
