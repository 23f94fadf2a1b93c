import csv, subprocess, sys, collections
rep = sys.argv[1]; want = sys.argv[2]
for skip in range(4):
    out = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass','--launch-skip',str(skip),'--launch-count','1'],capture_output=True,text=True).stdout
    lines = out.splitlines()
    rows = list(csv.reader(lines))
    # find function name rows
    fn = [r for r in rows if r and r[0]=='Function Name']
    if not fn or want not in fn[0][1]: continue
    agg = collections.OrderedDict()
    cur_file = None; H = None
    tot_i = tot_s = 0
    for r in rows:
        if not r: continue
        if r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
        if r[0] == 'Line No': H = r; ii = H.index('Instructions Executed'); it = H.index('Thread Instructions Executed'); isamp = H.index('# Samples'); continue
        if H is None or r[0] == 'Function Name': continue
        try:
            ln = int(r[0]); inst = int(r[ii]); thr = int(r[it]); samp = int(r[isamp])
        except: continue
        key = (cur_file, ln, r[1].strip()[:90])
        a = agg.setdefault(key, [0,0,0]); a[0]+=inst; a[1]+=thr; a[2]+=samp
        tot_i += inst; tot_s += samp
    if tot_i < 1e6: continue
    print('total inst', tot_i, 'samples', tot_s)
    for k,a in sorted(agg.items(), key=lambda kv: -kv[1][int(sys.argv[3]) if len(sys.argv)>3 else 2])[:45]:
        print(f"{k[0]:16s}:{k[1]:4d} inst {100*a[0]/tot_i:5.1f}% thr/inst {a[1]/max(a[0],1):5.1f} samples {100*a[2]/tot_s:5.1f}%  {k[2]}")
    break
