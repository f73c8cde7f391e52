#!/bin/bash
# registers / spills of every kernel variant (ptxas -v of the product's single translation unit)
cd "$(dirname "$0")/.."
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-ffp-contract=off -shared -Iinclude -Ipatchwork-plusplus_b200/csrc -cudart static -Xptxas=-v -o /tmp/pwpp_ptxas_report.so patchwork-plusplus_b200/csrc/pwpp_capi.cu 2>&1 | python -c "
import sys,re,subprocess
cur=None;rows=[]
for l in sys.stdin.read().split('\n'):
    m=re.search(r\"Compiling entry function '(\S+)'\",l)
    if m: cur=[m.group(1),'','']; rows.append(cur); continue
    if cur is None: continue
    if 'spill' in l: cur[1]=re.sub(r'.*: +','',l).strip()
    if 'Used' in l: cur[2]=re.sub(r'.*: +','',l).strip()
names=subprocess.run(['c++filt']+[r[0] for r in rows],capture_output=True,text=True).stdout.split('\n')
for r,n in zip(rows,names):
    n=re.sub(r'\(.*','',n).replace('void pwpp::','')
    print(f'{n:50s} {r[2][:34]:34s} | {r[1]}')
"
