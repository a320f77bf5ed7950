# usage: bash tools/ab_all.sh [variant names...]   -- default library first, then build_variants/lib_<name>.so
python tools/ab.py 64 512
for v in "$@"; do VQ_B200_LIB=build_variants/lib_$v.so python tools/ab.py 64 512; done
