"""Pipe pressure of a SASS range under the B200 rates measured by scripts/f32x2_probe.cu (profiles/r2_f32x2_probe.txt):
FFMA/FADD/FMUL 1 cycle, IMAD and the packed FFMA2/FADD2/FMUL2 2 cycles on the FMA pipe; ALU-pipe instructions 2 cycles;
MUFU 8 cycles (4 lanes per SMSP); one issue slot each.  usage: sass_pipe_count.py file.sass first_line last_line"""
import re, sys
FMA1 = {"FFMA", "FADD", "FMUL", "HFMA2", "HADD2", "HMUL2"}
FMA2 = {"IMAD", "FFMA2", "FADD2", "FMUL2", "IDP"}
ALU = {"IADD3", "LOP3", "SHF", "PRMT", "VIMNMX", "VIMNMX3", "ISETP", "SEL", "FSEL", "FMNMX", "I2FP", "LEA", "VIADD", "MOV", "FSETP",
       "PLOP3", "IABS", "BREV", "FLO", "POPC", "HMNMX2", "HSET2", "HSETP2", "CS2R", "F2FP", "I2IP", "VABSDIFF", "R2P", "P2R"}
XU = {"MUFU", "F2I", "I2F", "F2F"}
LSU = {"LDS", "STS", "LDG", "STG", "ATOMS", "LDC", "LDSM", "RED", "ATOM", "LDL", "STL"}
lines = open(sys.argv[1]).read().splitlines()[int(sys.argv[2]) - 1:int(sys.argv[3])]
cnt = {"fma": 0.0, "alu": 0.0, "xu": 0.0, "lsu": 0.0, "issue": 0, "other": {}}
for ln in lines:
    m = re.search(r"\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", ln)
    if not m:
        continue
    op = m.group(1)
    cnt["issue"] += 1
    if op in FMA1: cnt["fma"] += 1
    elif op in FMA2: cnt["fma"] += 2
    elif op in ALU: cnt["alu"] += 2
    elif op in XU: cnt["xu"] += 8
    elif op in LSU: cnt["lsu"] += 1
    else: cnt["other"][op] = cnt["other"].get(op, 0) + 1
print(cnt)
