import json, sys
d=json.load(open(sys.argv[1]))
tot=sum(r['ms'] for r in d['launches'])
for r in d['launches']: print(f"{r['launch']:22s} {r['kind']:8s} {r['ms']*1000:8.1f} us  {r['ms']/tot*100:5.1f}%")
print('total ms', tot)
