import re,sys,subprocess
# usage: ptxas_regs.py <log> [prefix ...]   -- registers / spills of the vl_dir_kernel instantiations
log=open(sys.argv[1]).read()
pref=tuple(sys.argv[2:]) or ('2, 0, 1','0, 0, 0')
cur=None; sp=''
for line in log.splitlines():
    m=re.search(r"Compiling entry function '(\S+)'",line)
    if m: cur=m.group(1); continue
    if 'spill' in line: sp=line.strip()
    m=re.search(r"Used (\d+) registers(.*)",line)
    if m and cur and 'vl_dir_kernel' in cur:
        name=subprocess.run(['c++filt',cur],capture_output=True,text=True).stdout.strip()
        mm=re.search(r"vl_dir_kernel<(.*?)>",name)
        if mm and mm.group(1).startswith(pref):
            print(mm.group(1), m.group(1), 'regs', sp, m.group(2)[:60])
