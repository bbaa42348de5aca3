python scripts/vertical_timeline.py 2>&1 | tail -3 | head -2
for v in e5 e6 e8; do
echo $v; WFB_LIB=wflow.jl_b200/csrc/_obj/variants/lib_$v.so python scripts/vertical_timeline.py 2>&1 | tail -3 | head -2
done
