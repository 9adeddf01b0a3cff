# usage: tl_sweep.sh "<cin> <cout> <res> <k>"  -- timeline under plan/debug knobs (PIPE PAIR NACC DBGMODE)
shape=${1:-"256 128 128 3"}; set -- $shape; cin=$1; cout=$2; res=$3; k=$4
for cfg in "0 1 0 0" "1 1 0 0" "0 1 0 7" "0 0 2 0" "0 0 2 7" "0 0 4 0" "1 0 4 0" "0 0 4 7" "0 1 0 3" "0 1 0 4"; do set -- $cfg; echo "== PIPE=$1 PAIR=$2 NACC=$3 DBGMODE=$4"; PASTA_B200_CONV_PIPE=$1 PASTA_B200_CONV_PAIR=$2 PASTA_B200_CONV_NACC=$3 PASTA_B200_CONV_DBGMODE=$4 python tools/conv_timeline.py --cin $cin --cout $cout --res $res --k $k | grep -E "CTAs|main loop|total|waiting|epilogue \(|fill"; done
