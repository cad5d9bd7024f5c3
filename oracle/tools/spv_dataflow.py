#!/usr/bin/env python3
"""Minimal SPIR-V walker used (at development time only) to read the floating-point dataflow of the
reference's shipped compute shaders, because the oracle pins that dataflow (DESIGN.md "arithmetic
contract").  It prints, per OpLine source line, every arithmetic instruction with its operands resolved
to short expressions.  Usage:  python oracle/tools/spv_dataflow.py /path/to/shader.comp.spv [first_line last_line]

Test infrastructure: nothing in the product, tests or bench imports this file; it reads a .spv given on the
command line and is not needed on the GPU box."""
import struct, sys

OPN = {  # opcode -> name (only what the hot-path shaders use)
    3: "Source", 5: "Name", 6: "MemberName", 8: "Line", 12: "ExtInst", 43: "Constant", 44: "ConstantComposite",
    50: "SpecConstant", 54: "Function", 56: "FunctionEnd", 59: "Variable", 61: "Load", 62: "Store",
    65: "AccessChain", 66: "InBoundsAccessChain", 71: "Decorate", 77: "VectorExtractDynamic",
    79: "VectorShuffle", 80: "CompositeConstruct", 81: "CompositeExtract", 82: "CompositeInsert",
    84: "Transpose", 109: "ConvertFToU", 110: "ConvertFToS", 111: "ConvertSToF", 112: "ConvertUToF",
    124: "Bitcast", 126: "SNegate", 127: "FNegate", 128: "IAdd", 129: "FAdd", 130: "ISub", 131: "FSub",
    132: "IMul", 133: "FMul", 134: "UDiv", 135: "SDiv", 136: "FDiv", 137: "UMod", 142: "VectorTimesScalar",
    143: "MatrixTimesScalar", 144: "VectorTimesMatrix", 145: "MatrixTimesVector", 146: "MatrixTimesMatrix",
    148: "Dot", 164: "LogicalEqual", 166: "LogicalOr", 167: "LogicalAnd", 168: "LogicalNot", 169: "Select",
    170: "IEqual", 171: "INotEqual", 172: "UGreaterThan", 174: "UGreaterThanEqual", 176: "ULessThan",
    178: "ULessThanEqual", 180: "FOrdEqual", 184: "FOrdLessThan", 186: "FOrdGreaterThan",
    188: "FOrdLessThanEqual", 190: "FOrdGreaterThanEqual", 185: "FUnordLessThan", 187: "FUnordGreaterThan",
    194: "ShiftRightLogical", 196: "ShiftLeftLogical", 197: "BitwiseOr", 199: "BitwiseAnd", 245: "Phi",
    246: "LoopMerge", 247: "SelectionMerge", 248: "Label", 249: "Branch", 250: "BranchConditional",
    251: "Switch", 253: "Return", 87: "SampledImage", 88: "ImageSampleExplicitLod", 86: "Image",
    95: "ImageFetch", 99: "ImageWrite", 103: "ImageQuerySizeLod", 234: "AtomicIAdd", 240: "AtomicUMax",
    242: "AtomicOr", 339: "GroupNonUniformBallot", 333: "GroupNonUniformElect", 4472: "SubgroupBallotKHR",
}
GLSL = {4: "FAbs", 8: "Floor", 26: "Pow", 30: "Log2", 31: "Sqrt", 37: "FMin", 40: "FMax", 38: "UMin",
        41: "UMax", 43: "FClamp", 44: "UClamp", 46: "FMix", 50: "Fma", 66: "Length", 67: "Distance",
        69: "Normalize", 79: "NMin", 80: "NMax", 81: "NClamp"}

def main():
    path = sys.argv[1]
    lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10**9
    w = struct.unpack("<%dI" % (len(open(path, "rb").read()) // 4), open(path, "rb").read())
    assert w[0] == 0x07230203
    i, line, consts, names = 5, 0, {}, {}
    def nm(x):
        if x in consts: return "c(%s)" % (consts[x],)
        return names.get(x, "%%%d" % x)
    while i < len(w):
        op, n = w[i] & 0xFFFF, w[i] >> 16
        a = w[i + 1:i + n]
        if op == 5:
            s = b"".join(struct.pack("<I", x) for x in a[1:]).split(b"\0")[0].decode()
            if s: names[a[0]] = s + "#%d" % a[0]
        elif op == 43 and len(a) == 3:
            f = struct.unpack("<f", struct.pack("<I", a[2]))[0]
            consts[a[1]] = "%r|0x%08x" % (f, a[2])
        elif op == 8:
            line = a[1]
        elif lo <= line <= hi and op in OPN and op not in (3, 5, 6, 8, 71, 43, 44, 59, 248, 247, 246, 253, 56):
            name = OPN[op]
            if op == 12:
                print("L%-4d %%%d = %s %s" % (line, a[1], GLSL.get(a[3], "ext%d" % a[3]), " ".join(nm(x) for x in a[4:])))
            elif op in (62, 249, 250, 251, 253, 99):
                print("L%-4d %s %s" % (line, name, " ".join(nm(x) for x in a)))
            elif op in (79, 81, 82):
                print("L%-4d %%%d = %s %s" % (line, a[1], name, " ".join(nm(x) if k < (2 if op == 79 else (2 if op == 82 else 1)) else str(x) for k, x in enumerate(a[2:]))))
            else:
                print("L%-4d %%%d = %s %s" % (line, a[1], name, " ".join(nm(x) for x in a[2:])))
        i += n

if __name__ == "__main__":
    main()
