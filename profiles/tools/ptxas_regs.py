import sys,re
L=sys.stdin.read().split('\n')
name=None; sp=''
for l in L:
    m=re.search(r"function '_ZN5slode(\d+)(\w+?)ILi(\d+)ELi(\d+)ELi(\d+)(?:ELi(\d+))?",l)
    if m: name=(m.group(2)[:18],m.group(3),m.group(4),m.group(5),m.group(6))
    if 'spill' in l: sp=l.strip()
    if 'Used' in l and name: print(name, l.strip().split(',')[0], '|', sp); name=None
    if 'error' in l: print(l)
