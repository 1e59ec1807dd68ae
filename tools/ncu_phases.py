#!/usr/bin/env python3
"""Aggregate tools/ncu_lines.py output of k_l2 into phases: python tools/ncu_phases.py lines.txt"""
import re, sys
K=[(0,700,'k_l2_cross/other'),(700,782,'cross'),(782,830,'setup'),(830,863,'ray'),(863,917,'satA'),(917,951,'satB1'),(951,968,'satB2'),(968,1005,'out')]
M=[(0,36,'m:axis_sep(sat)'),(36,100,'m:tri_box'),(100,132,'m:row_setup'),(132,157,'m:row_values/clip'),(157,190,'m:row_test'),(190,225,'m:plane'),(225,250,'m:ray_setup'),(250,262,'m:ray_column'),(262,271,'m:ray_cell'),(271,300,'m:z_run'),(300,400,'m:other')]
tot={}
for l in open(sys.argv[1]):
    m=re.match(r'\s*([\d.]+)% inst\s+([\d.]+)% stall\s+lanes\s+([\d.]+)\s+(\S+):(\d+)\s+(.*)',l)
    if not m: continue
    pi,ps,lanes,f,ln=float(m[1]),float(m[2]),float(m[3]),m[4],int(m[5])
    tab = M if f=='gpv_math.h' else K if f=='gpv_kernels.cuh' else None
    key=f
    if tab:
        for a,b,n in tab:
            if a<=ln<b: key=n
    v=tot.setdefault(key,[0,0,0]); v[0]+=pi; v[1]+=ps; v[2]+=pi*lanes
for k,v in sorted(tot.items(), key=lambda kv:-kv[1][0]): print('%-28s %5.1f%% inst %5.1f%% stall  lanes %4.1f'%(k,v[0],v[1],v[2]/max(v[0],1e-9)))
