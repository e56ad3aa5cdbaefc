# CTA size of the traversal kernels: 128 threads x 8 per SM (production) against 64 x 16 and 32 x 32 (builds made with
# make "NVFLAGS+=-DMB200_BLOCK=64" ...; the libraries are swapped in place on the GPU box)
cp mallie_b200/libmallie_b200.so /tmp/lib128.so
for b in 128 64 32; do
  if [ $b = 128 ]; then cp /tmp/lib128.so mallie_b200/libmallie_b200.so; else cp mallie_b200/libmallie_b200.so.b$b mallie_b200/libmallie_b200.so; fi
  echo "== block $b"
  python tools/ab_frame.py 2>&1 | tail -1
  AB_N=8 python tools/ab_band.py 2>&1 | tail -1
done
cp /tmp/lib128.so mallie_b200/libmallie_b200.so
