import csv, subprocess, sys, io
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(out)))
h=rows[1]; body=rows[2:]
ix={k:i for i,k in enumerate(h)}
tot=sum(int(r[ix['# Samples']]) for r in body)
print('total samples', tot, 'instr', len(body))
top=sorted(range(len(body)), key=lambda i:-int(body[i][ix['# Samples']]))[:int(sys.argv[2]) if len(sys.argv)>2 else 30]
for i in sorted(top):
    r=body[i]
    st={k:int(r[ix[k]]) for k in h if k.startswith('stall_') and 'Not' not in k and r[ix[k]] not in ('','0')}
    st=dict(sorted(st.items(), key=lambda kv:-kv[1])[:3])
    print(i, r[ix['Source']].strip()[:64], r[ix['# Samples']], r[ix['L1 Wavefronts Shared']], r[ix['L1 Wavefronts Shared Ideal']], st)
acc=0
for i,r in enumerate(body):
    acc+=int(r[ix['# Samples']])
    if 'BAR.SYNC' in r[ix['Source']] or i==len(body)-1:
        print('--- up to instr', i, 'samples', acc, '(%.1f%%)'%(100*acc/tot)); acc=0
