import csv, subprocess, sys, collections
rep = sys.argv[1]; want = sys.argv[2]
for skip in range(4):
    out = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass','--launch-skip',str(skip),'--launch-count','1'],capture_output=True,text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    fn = [r for r in rows if r and r[0]=='Function Name']
    if not fn or want not in fn[0][1]: continue
    agg = collections.OrderedDict(); cur_file=None; H=None; tot=0
    for r in rows:
        if not r: continue
        if r[0]=='File Path': cur_file=r[1].split('/')[-1]; continue
        if r[0]=='Line No': H=r; ii=H.index('Instructions Executed'); it=H.index('Thread Instructions Executed'); continue
        if H is None or r[0]=='Function Name': continue
        try: ln=int(r[0]); inst=int(r[ii]); thr=int(r[it])
        except: continue
        a=agg.setdefault((cur_file,ln),[0,0,r[1].strip()[:70]]); a[0]+=inst; a[1]+=thr; tot+=inst
    if tot<1e6: continue
    print('total', tot)
    for (f,ln),a in sorted(agg.items()):
        if a[0] > tot*0.003: print(f"{f:14s}:{ln:4d} {100*a[0]/tot:5.1f}% thr {a[1]/max(a[0],1):5.1f}  {a[2]}")
    break
