import csv,sys,collections
for tag in sys.argv[1:]:
    rows=list(csv.DictReader(open(f"gpurun_out/layers_{tag}.csv")))
    tot=sum(float(r['ms']) for r in rows)
    g=[r for r in rows if r['family']=='dp_gemm_tc']
    by=collections.defaultdict(lambda:[0,0.0,0.0])
    for r in g:
        b=by[r['label']]; b[0]+=1; b[1]+=float(r['ms']); b[2]+=float(r['gflop'])
    print(tag,"total %.2f gemm %.3f ms"%(tot,sum(float(r['ms']) for r in g)))
    for k,(n,ms,gf) in sorted(by.items(), key=lambda x:-x[1][1]):
        print("   %2d x %-45s %.3f ms  %.0f TF/s"%(n,k,ms,gf/ms if ms else 0))
