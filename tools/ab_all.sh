python tools/ab.py 64 512
for v in abl_norerank abl_noout abl_noq ow8 ncg4 ncg1; do VQ_B200_LIB=build_variants/lib_$v.so python tools/ab.py 64 512; done
