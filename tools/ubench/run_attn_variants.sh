# Knock-out timing of the attention kernel's softmax phases (debug builds, wrong results).  Build the variants with
# -DATT_DBG_NOMAX / NOSCALE / NOSUM / NOEXP into tools/ubench/variants/lib_<name>.so first.
cp walkgpt_b200/libwalkgpt_b200.so /tmp/orig.so
for v in NOMAX ALL3 ALL4; do
  if [ $v = orig ]; then cp /tmp/orig.so walkgpt_b200/libwalkgpt_b200.so; else cp tools/ubench/variants/lib_$v.so walkgpt_b200/libwalkgpt_b200.so; fi
  echo "== $v"; T=1024 timeout 120 python tools/time_attn.py 2>&1 | tail -1
done
cp /tmp/orig.so walkgpt_b200/libwalkgpt_b200.so
