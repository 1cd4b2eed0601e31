"""Per-launch table (time, DRAM bytes, algorithmic FLOPs/bytes) from an ncu --csv launch list of one bf16 decode."""
import csv, sys, json
f=sys.argv[1]; frames=int(sys.argv[2]) if len(sys.argv)>2 else 3744
lines=[l for l in open(f) if l.startswith('"')]
per={}
for r in csv.DictReader(lines):
    per.setdefault(int(r['ID']),{})[r['Metric Name']]=(float(r['Metric Value'].replace(',','')), r['Metric Unit'])
ids=sorted(per)
names=['conv_pre']
for s in range(6):
    names.append(f's{s} ups')
    for k in [3,7,11]:
        for d in [1,3,5]:
            names.append(f's{s} k{k} d{d} A'); names.append(f's{s} k{k} d{d} B')
C=[768,384,192,96,48,24]; R=[4,16,64,256,512,1024]
def conv(v,u):
    return v*{'ns':1e-3,'us':1,'usecond':1,'nsecond':1e-3,'ms':1e3,'msecond':1e3,'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}.get(u,1)
rows=[]; tot_t=0; tot_tr=0; tot_alg=0; tot_fl=0
for n,i in zip(names,ids):
    m=per[i]
    t=conv(*m['gpu__time_duration.sum'])
    tr=conv(*m.get('dram__bytes_read.sum',(0,'byte')))+conv(*m.get('dram__bytes_write.sum',(0,'byte')))
    p=n.split(); fl=alg=0
    if len(p)==4:
        s=int(p[0][1]); k=int(p[1][1:]); el=C[s]*R[s]*frames
        fl=2*C[s]*k*el; alg=el*2*(3 if p[3]=='B' else 2)+C[s]*C[s]*k*2
        if p[3]=='B' and p[2]=='d5' : alg+= el*2*(1 if k>3 else 0)
        tot_t+=t; tot_tr+=tr; tot_alg+=alg; tot_fl+=fl
    rows.append((n,t,tr,fl,alg))
print(f"{'launch':14s} {'us':>8s} {'TFLOP/s':>8s} {'dram MB':>8s} {'alg MB':>8s} {'GB/s(alg)':>9s}")
for n,t,tr,fl,alg in rows:
    print(f"{n:14s} {t:8.1f} {fl/t/1e6 if fl else 0:8.1f} {tr/1e6:8.1f} {alg/1e6:8.1f} {alg/t/1e3 if alg else 0:9.1f}")
print(json.dumps({"amp_launches":108,"amp_time_us":tot_t,"amp_tflops":tot_fl/tot_t/1e6,"amp_dram_bytes_per_launch":tot_tr/108,"amp_alg_bytes_per_launch":tot_alg/108,"amp_flops_per_launch":tot_fl/108}))
