import sys, json
sys.path.insert(0,'/root/repo/tools'); sys.path.insert(0,'/root/repo')
from conv_sweep import run
name=sys.argv[1]; n=int(sys.argv[2]); ov=json.loads(sys.argv[3]) if len(sys.argv)>3 else None
r,by,fl=run(name,n,ov,2,'fp16'); print(name,ov,r['ms'],r['plan'])
