"""oracle/pyrlp.py -- TEST INFRASTRUCTURE. Minimal RLP list splitter for oracle/pytrie.py."""


def item_len(buf: bytes, pos: int) -> int:
    b = buf[pos]
    if b < 0x80:
        return 1
    if b < 0xB8:
        return 1 + b - 0x80
    if b < 0xC0:
        ll = b - 0xB7
        return 1 + ll + int.from_bytes(buf[pos + 1:pos + 1 + ll], "big")
    if b < 0xF8:
        return 1 + b - 0xC0
    ll = b - 0xF7
    return 1 + ll + int.from_bytes(buf[pos + 1:pos + 1 + ll], "big")


def split_list(enc: bytes):
    """Items (with their headers) of a well-formed RLP list, or None if enc is not a list."""
    if not enc or enc[0] < 0xC0:
        return None
    b = enc[0]
    if b < 0xF8:
        pos, end = 1, 1 + b - 0xC0
    else:
        ll = b - 0xF7
        pos = 1 + ll
        end = pos + int.from_bytes(enc[1:1 + ll], "big")
    out = []
    while pos < end:
        n = item_len(enc, pos)
        out.append(enc[pos:pos + n])
        pos += n
    return out
