import sys,csv
rows=[r for r in csv.reader(sys.stdin) if len(r)>12 and 'tn_pair' in r[4]]
rd=sum(float(r[-1]) for r in rows if r[-3]=='dram__bytes_read.sum'); t=sum(float(r[-1]) for r in rows if r[-3]=='gpu__time_duration.sum')
n=sum(1 for r in rows if r[-3]=='gpu__time_duration.sum')
print(f"launches {n} dram_read {rd/1e9:.2f} GB time {t/1e3:.0f} us")
