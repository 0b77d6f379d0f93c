"""
Minimal stand-in for the ``bitarray`` package (absent from this image), just
enough for /root/reference/vc2_conformance/bitstream/{io,vc2_fixeddicts}.py and
test_cases/decoder/pictures.py.
TEST INFRASTRUCTURE ONLY (golden-vector generation; see oracle/ref_import.py).
"""


class bitarray(object):
    def __init__(self, init=()):
        if isinstance(init, str):
            self._bits = [1 if ch == "1" else 0 for ch in init]
        elif isinstance(init, int):
            self._bits = [0] * init
        else:
            self._bits = [1 if b else 0 for b in init]

    def append(self, bit):
        self._bits.append(1 if bit else 0)

    def extend(self, other):
        for b in other:
            self.append(b)

    def __add__(self, other):
        return bitarray(self._bits + list(other))

    def __len__(self):
        return len(self._bits)

    def __iter__(self):
        return iter(self._bits)

    def __getitem__(self, idx):
        if isinstance(idx, slice):
            return bitarray(self._bits[idx])
        return self._bits[idx]

    def __eq__(self, other):
        return isinstance(other, bitarray) and self._bits == other._bits

    def __ne__(self, other):
        return not self == other

    def __hash__(self):
        return hash(tuple(self._bits))

    def frombytes(self, data):
        """bitarray.frombytes: appends the bits of `data`, most significant bit of each byte first
        (the real package's default big-endian bit order, which the reference relies on)."""
        for byte in bytearray(data):
            for k in range(7, -1, -1):
                self._bits.append((byte >> k) & 1)

    def to01(self):
        return "".join(str(b) for b in self._bits)

    def tobytes(self):
        bits = self._bits + [0] * ((-len(self._bits)) % 8)
        out = bytearray()
        for i in range(0, len(bits), 8):
            v = 0
            for b in bits[i : i + 8]:
                v = (v << 1) | b
            out.append(v)
        return bytes(out)

    def __repr__(self):
        return "bitarray('{}')".format(self.to01())
