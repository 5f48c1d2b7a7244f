export RB_TMP=/tmp/rbt
NSENS=512 python tools/dev_prof.py > /dev/null 2>&1
for so in pyradiance_b200/librb200.so variants_nodiff.so; do
  echo "== $so"; RB200_LIBRARY=$PWD/$so python tools/walltest.py 2>&1 | tail -3
done
