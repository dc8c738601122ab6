"""write a .safetensors file with numpy (huggingface/safetensors format: u64 LE header length | JSON header | data)"""
import json

import numpy as np

DT = {"BF16": np.uint16, "F16": np.float16, "F32": np.float32, "I32": np.int32}


def write_safetensors(path, tensors, metadata=None):
    """tensors: list of (name, dtype string, array) -- a BF16 array is given as uint16 bit patterns"""
    header, blobs, off = {}, [], 0
    if metadata:
        header["__metadata__"] = metadata
    for name, dt, a in tensors:
        a = np.ascontiguousarray(a, dtype=DT[dt])
        header[name] = {"dtype": dt, "shape": list(a.shape), "data_offsets": [off, off + a.nbytes]}
        blobs.append(a.tobytes())
        off += a.nbytes
    text = json.dumps(header, separators=(",", ":")).encode()
    text += b" " * ((8 - len(text) % 8) % 8)
    with open(path, "wb") as f:
        f.write(len(text).to_bytes(8, "little"))
        f.write(text)
        for b in blobs:
            f.write(b)


def write_kun_reference_style(path, tensors, config, moments=1):
    """tensors: [(name, dtype name, shape, szData, szGama, blob bytes)]; the config entry registered after the tensors, as insertJS does"""
    header, off, blobs = {"__metadata__": {"format": "pt", "writer": "koifish"}}, 0, []
    for name, dt, shape, szd, szg, blob in tensors:
        payload = bytes(blob) * moments
        header[name] = {"dtype": dt, "shape": list(shape), "data_offsets": [off, off + len(payload)], "loAB": 0, "szGama": szg, "szData": szd}
        off += len(payload)
        blobs.append(payload)
    if config is not None:
        import msgpack
        mp = msgpack.packb(config, use_bin_type=True)
        header["__koifish__config__"] = {"dtype": "U8", "shape": [len(mp)], "data_offsets": [off, off + len(mp)], "loAB": 0, "szGama": 0, "szData": 0}
        blobs.append(mp)
    text = json.dumps(header).encode()
    with open(path, "wb") as f:
        f.write(len(text).to_bytes(8, "little") + text + b"".join(blobs))
