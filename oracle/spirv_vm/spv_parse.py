"""SPIR-V binary -> Python structures (test infrastructure; see spirv_vm.py)."""
import struct


class Inst:
    __slots__ = ("op", "words", "result", "rtype", "line", "index")

    def __init__(self, op, words, line, index=0):
        self.op, self.words, self.line, self.index = op, words, line, index
        self.result = self.rtype = None

    def __repr__(self):
        return "Inst(op=%d, %s)" % (self.op, list(self.words))


def string_at(words, i):
    b = b"".join(struct.pack("<I", x) for x in words[i:])
    s = b.split(b"\0")[0]
    return s.decode(), i + len(s) // 4 + 1


# opcodes with (result type, result id) as first two operands / with only a result id
HAS_TYPE_AND_RESULT = {1, 12, 41, 42, 43, 44, 45, 46, 48, 49, 50, 51, 52, 54, 55, 57, 59, 61, 65, 66, 67, 77, 78, 79, 80, 81, 82, 83, 84,
                       86, 87, 88, 89, 95, 100, 103, 104, 109, 110, 111, 112, 113, 114, 115, 116, 124, 126, 127, 128, 129, 130, 131, 132,
                       133, 134, 135, 136, 137, 138, 139, 140, 141, 142, 143, 144, 145, 146, 147, 148, 154, 155, 156, 157, 164, 165, 166,
                       167, 168, 169, 170, 171, 172, 173, 174, 175, 176, 177, 178, 179, 180, 181, 182, 183, 184, 185, 186, 187, 188, 189,
                       190, 191, 194, 195, 196, 197, 198, 199, 200, 201, 202, 203, 204, 205, 227, 229, 230, 231, 232, 233, 234, 235, 236,
                       237, 238, 239, 240, 241, 242, 245, 333, 334, 335, 336, 337, 338, 339, 340, 341, 342, 343, 344, 345, 400}
HAS_RESULT_ONLY = {7, 11, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 248}


class Module:
    def __init__(self, path):
        data = open(path, "rb").read()
        w = struct.unpack("<%dI" % (len(data) // 4), data)
        assert w[0] == 0x07230203, "not SPIR-V"
        self.version, self.bound = w[1], w[3]
        self.insts = []
        self.names, self.member_names = {}, {}
        self.decor, self.member_decor = {}, {}      # id -> {decoration: [operands]} ; (id, member) -> {...}
        self.defs = {}                               # result id -> Inst
        self.entry = None
        self.exec_modes = {}
        i, line = 5, 0
        while i < len(w):
            op, n = w[i] & 0xFFFF, w[i] >> 16
            a = w[i + 1:i + n]
            if op == 8:
                line = a[1]
            elif op == 317:
                line = 0
            inst = Inst(op, a, line, len(self.insts))
            if op in HAS_TYPE_AND_RESULT:
                inst.rtype, inst.result = a[0], a[1]
            elif op in HAS_RESULT_ONLY:
                inst.result = a[0]
            if inst.result is not None:
                self.defs[inst.result] = inst
            if op == 5:
                self.names[a[0]] = string_at(a, 1)[0]
            elif op == 6:
                self.member_names[(a[0], a[1])] = string_at(a, 2)[0]
            elif op == 71:
                self.decor.setdefault(a[0], {})[a[1]] = list(a[2:])
            elif op == 72:
                self.member_decor.setdefault((a[0], a[1]), {})[a[2]] = list(a[3:])
            elif op == 15:
                name, j = string_at(a, 2)
                self.entry = {"model": a[0], "id": a[1], "name": name, "interface": list(a[j:])}
            elif op == 16:
                self.exec_modes[a[1]] = list(a[2:])
            self.insts.append(inst)
            i += n
