import json, sys
rows = [json.loads(l) for l in open(sys.argv[1])]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
tot = sum(r['ms_total'] for r in rows)
fl = sum(2.0 * r['M'] * r['N'] * r['K'] * r['launches'] for r in rows)
print("GEMM total %.3f ms/step, %.1f TFLOP/s aggregate" % (tot / steps, fl / tot / 1e9))
for r in rows[:int(sys.argv[3]) if len(sys.argv) > 3 else 26]:
    print("%8d %6d %8d a%d b%d x%-3d %7.3f ms/step %6.3f ms avg %7.1f TF" % (r['M'], r['N'], r['K'], r['a_mn'], r['b_mn'], r['launches'] // steps, r['ms_total'] / steps, r['ms_avg'], r['tflops']))
