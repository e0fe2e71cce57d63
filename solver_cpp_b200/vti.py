"""VTK ImageData (.vti) grid files as DAMASK writes and reads them: the reference's geometry input
(/root/reference/src/grid/VTI.f90:146-499 reader: cells from WholeExtent, size = Spacing * cells, Origin, and the cell
data array `material`, base64 + optional vtkZLibDataCompressor blocks; /root/reference/python/damask/_grid.py:save).
Host-side input plumbing of the drop-in; numpy + zlib + base64 only."""
import base64
import re
import zlib

import numpy as np

_DTYPES = {'Int8': 'i1', 'UInt8': 'u1', 'Int16': '<i2', 'UInt16': '<u2', 'Int32': '<i4', 'UInt32': '<u4', 'Int64': '<i8', 'UInt64': '<u8',
           'Float32': '<f4', 'Float64': '<f8'}


def _attr(tag_text, name):
    m = re.search(r'\b%s\s*=\s*"([^"]*)"' % re.escape(name), tag_text)
    if not m:
        raise ValueError('VTI: attribute %s missing in <%s>' % (name, tag_text.split()[0]))
    return m.group(1)


def _b64_len(n_bytes):
    return 4 * ((n_bytes + 2) // 3)


def _decode_binary(text, dtype, header_type, compressed):
    """payload of a format="binary" DataArray: [header][data], each base64-encoded on its own"""
    text = ''.join(text.split())
    hdt = np.dtype(_DTYPES[header_type])
    if not compressed:
        raw = base64.b64decode(text)
        n = int(np.frombuffer(raw[:hdt.itemsize], hdt)[0])
        return np.frombuffer(raw[hdt.itemsize:hdt.itemsize + n], dtype)
    first = base64.b64decode(text[:_b64_len(3 * hdt.itemsize)])
    n_blocks = int(np.frombuffer(first[:hdt.itemsize], hdt)[0])
    hlen = _b64_len((3 + n_blocks) * hdt.itemsize)
    head = np.frombuffer(base64.b64decode(text[:hlen])[:(3 + n_blocks) * hdt.itemsize], hdt)
    sizes = [int(s) for s in head[3:3 + n_blocks]]
    data = base64.b64decode(text[hlen:])
    out, pos = [], 0
    for s in sizes:
        out.append(zlib.decompress(data[pos:pos + s]))
        pos += s
    return np.frombuffer(b''.join(out), dtype)


def read_vti(path, dataset='material'):
    """-> dict(cells (3,), size (3,), origin (3,), material [nz, ny, nx] as stored in the file (0-based indices into
    material.yaml's `material` list; the Fortran reader adds 1, VTI.f90:474))"""
    txt = open(path, 'r').read()
    vtk = re.search(r'<VTKFile\b[^>]*>', txt).group(0)
    if _attr(vtk, 'type') != 'ImageData':
        raise ValueError('VTI: not an ImageData file')
    if _attr(vtk, 'byte_order') != 'LittleEndian':
        raise ValueError('VTI: only little-endian files are supported')
    header_type = re.search(r'header_type\s*=\s*"([^"]*)"', vtk)
    header_type = header_type.group(1) if header_type else 'UInt32'
    compressed = 'compressor' in vtk
    if compressed and _attr(vtk, 'compressor') != 'vtkZLibDataCompressor':
        raise ValueError('VTI: unsupported compressor')
    img = re.search(r'<ImageData\b[^>]*>', txt).group(0)
    ext = [int(v) for v in _attr(img, 'WholeExtent').split()]
    cells = np.array([ext[1] - ext[0], ext[3] - ext[2], ext[5] - ext[4]])
    spacing = np.array([float(v) for v in _attr(img, 'Spacing').split()])
    origin = np.array([float(v) for v in _attr(img, 'Origin').split()])
    cd = re.search(r'<CellData\b[^>]*>(.*?)</CellData>', txt, re.S)
    if not cd:
        raise ValueError('VTI: no CellData')
    for m in re.finditer(r'<DataArray\b([^>]*)>(.*?)</DataArray>', cd.group(1), re.S):
        tag = '<DataArray' + m.group(1) + '>'
        if _attr(tag, 'Name') != dataset:
            continue
        dtype = _DTYPES[_attr(tag, 'type')]
        fmt = _attr(tag, 'format')
        if fmt == 'binary':
            a = _decode_binary(m.group(2), dtype, header_type, compressed)
        elif fmt == 'ascii':
            a = np.array(m.group(2).split(), dtype=np.dtype(dtype))
        else:
            raise ValueError('VTI: unsupported DataArray format %s' % fmt)
        if a.size != int(np.prod(cells)):
            raise ValueError('VTI: dataset %s has %d entries, grid has %d cells' % (dataset, a.size, int(np.prod(cells))))
        return dict(cells=cells, size=spacing * cells, origin=origin,
                    material=np.ascontiguousarray(a.reshape(cells[2], cells[1], cells[0])).astype(np.int64))
    raise ValueError('VTI: cell dataset %s not found' % dataset)


def write_vti(path, material, size, origin=(0.0, 0.0, 0.0), compress=True, block=32768):
    """material [nz, ny, nx] (0-based) -> a file DAMASK's reader accepts (Int64 `material`, UInt32 headers, zlib blocks)"""
    m = np.ascontiguousarray(material, dtype='<i8')
    nz, ny, nx = m.shape
    raw = m.tobytes()
    if compress:
        blocks = [raw[i:i + block] for i in range(0, len(raw), block)] or [b'']
        comp = [zlib.compress(b) for b in blocks]
        last = len(blocks[-1]) if len(blocks[-1]) != block else 0
        head = np.array([len(blocks), block, last] + [len(c) for c in comp], dtype='<u4').tobytes()
        payload = base64.b64encode(head).decode() + base64.b64encode(b''.join(comp)).decode()
    else:
        payload = base64.b64encode(np.array([len(raw)], dtype='<u4').tobytes() + raw).decode()
    spacing = np.asarray(size, float) / np.array([nx, ny, nz])
    with open(path, 'w') as f:
        f.write('<?xml version="1.0"?>\n<VTKFile type="ImageData" version="0.1" byte_order="LittleEndian" header_type="UInt32"%s>\n'
                % (' compressor="vtkZLibDataCompressor"' if compress else ''))
        f.write('  <ImageData WholeExtent="0 %d 0 %d 0 %d" Origin="%s" Spacing="%s" Direction="1 0 0 0 1 0 0 0 1">\n'
                % (nx, ny, nz, ' '.join(repr(float(v)) for v in origin), ' '.join(repr(float(v)) for v in spacing)))
        f.write('  <Piece Extent="0 %d 0 %d 0 %d">\n    <PointData>\n    </PointData>\n    <CellData>\n' % (nx, ny, nz))
        f.write('      <DataArray type="Int64" Name="material" format="binary" RangeMin="%d" RangeMax="%d">\n        %s\n      </DataArray>\n'
                % (int(m.min()), int(m.max()), payload))
        f.write('    </CellData>\n  </Piece>\n  </ImageData>\n</VTKFile>\n')
