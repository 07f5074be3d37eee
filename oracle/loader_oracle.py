"""TEST INFRASTRUCTURE -- CPU restatement of the reference's *loader* (BAM + GTF -> uint16 score
matrix), used only to build the config-1 fixture (BASELINE.json configs[0]) from the files the
reference bundles, so that the EM path can be pinned against ``stellarscope/data/
telescope_report.tsv`` (the one golden the reference ships for this path).

The reference needs pysam + intervaltree for this (absent here), so the few steps are restated in
plain Python, each citing what it follows (paths relative to the reference checkout,
``stellarscope/``):

* BGZF / BAM decoding: SAM/BAM format specification v1 (what pysam/htslib implement).
* ``fetch_fragments_seq`` / ``pair_bundle`` / ``readkey`` / ``matekey``: utils/alignment.py:34-90, 171-214
* ``AlignedPair.refblocks / alnlen / alnscore``: utils/calignment.pyx:96-111; ``merge_blocks``: utils/helpers.py:92-122
* annotation: annotation/_intervaltree.py:34-84 (per-locus exon merge, intervals [start, end+1)),
  annotation/utils.py:9-46 (``overlap_length``, ``merge_neighbors``), GTF parsing :48-118
* ``Assigner`` threshold mode: utils/model.py:519-537
* ``process_overlap_frag``: utils/model.py:51-84
* ``_load_sequential`` / ``mapping_to_matrix``: utils/model.py:958-978, 996-1146; minAS: utils/statistics.py:230-231

Nothing here is imported by the product.
"""
from __future__ import annotations

import gzip
import re
import struct
from collections import Counter, defaultdict

import numpy as np


# ---------------------------------------------------------------------------------------------
# BAM
# ---------------------------------------------------------------------------------------------
class Aln:
    __slots__ = ('qname', 'flag', 'ref_id', 'pos', 'next_ref_id', 'next_pos', 'tlen', 'cigar', 'tags')

    is_paired = property(lambda s: bool(s.flag & 0x1))
    is_proper_pair = property(lambda s: bool(s.flag & 0x2))
    is_unmapped = property(lambda s: bool(s.flag & 0x4))
    is_reverse = property(lambda s: bool(s.flag & 0x10))
    is_read1 = property(lambda s: bool(s.flag & 0x40))

    def get_blocks(self):
        """pysam ``AlignedSegment.get_blocks``: gapless aligned blocks (M, =, X), 0-based half-open."""
        out, p = [], self.pos
        for op, ln in self.cigar:
            if op in (0, 7, 8):          # M = X
                out.append((p, p + ln))
                p += ln
            elif op in (2, 3):           # D N
                p += ln
        return out


_TAG_FMT = {'c': 'b', 'C': 'B', 's': 'h', 'S': 'H', 'i': 'i', 'I': 'I', 'f': 'f'}


def _parse_tags(buf, off, end):
    tags = {}
    while off < end:
        tag = buf[off:off + 2].decode()
        typ = chr(buf[off + 2])
        off += 3
        if typ == 'A':
            tags[tag] = chr(buf[off])
            off += 1
        elif typ in _TAG_FMT:
            f = _TAG_FMT[typ]
            n = struct.calcsize(f)
            tags[tag] = struct.unpack_from('<' + f, buf, off)[0]
            off += n
        elif typ in 'ZH':
            z = buf.index(b'\0', off)
            tags[tag] = buf[off:z].decode()
            off = z + 1
        elif typ == 'B':
            sub = chr(buf[off])
            cnt = struct.unpack_from('<i', buf, off + 1)[0]
            n = struct.calcsize(_TAG_FMT[sub])
            off += 5 + n * cnt
        else:
            raise ValueError(f'unknown BAM tag type {typ}')
    return tags


def read_bam(path):
    """(reference names, list of Aln in file order)."""
    buf = gzip.open(path, 'rb').read()          # BGZF = concatenated gzip members
    assert buf[:4] == b'BAM\1'
    l_text = struct.unpack_from('<i', buf, 4)[0]
    off = 8 + l_text
    n_ref = struct.unpack_from('<i', buf, off)[0]
    off += 4
    refs = []
    for _ in range(n_ref):
        l_name = struct.unpack_from('<i', buf, off)[0]
        refs.append(buf[off + 4:off + 4 + l_name - 1].decode())
        off += 4 + l_name + 4
    alns = []
    while off < len(buf):
        (block_size, ref_id, pos, l_read_name, _mapq, _bin, n_cigar, flag, l_seq, next_ref, next_pos,
         tlen) = struct.unpack_from('<iiiBBHHHiiii', buf, off)
        rec_end = off + 4 + block_size
        p = off + 36
        a = Aln()
        a.qname = buf[p:p + l_read_name - 1].decode()
        p += l_read_name
        a.cigar = [(c & 0xf, c >> 4) for c in struct.unpack_from(f'<{n_cigar}I', buf, p)]
        p += 4 * n_cigar + (l_seq + 1) // 2 + l_seq
        a.flag, a.ref_id, a.pos, a.next_ref_id, a.next_pos, a.tlen = flag, ref_id, pos, next_ref, next_pos, tlen
        a.tags = _parse_tags(buf, p, rec_end)
        alns.append(a)
        off = rec_end
    return refs, alns


# ---------------------------------------------------------------------------------------------
# fragments (utils/alignment.py, utils/calignment.pyx)
# ---------------------------------------------------------------------------------------------
def merge_blocks(ivs, dist=0):
    """utils/helpers.py:92-122"""
    if len(ivs) <= 1:
        return ivs
    ivs = sorted(ivs, key=lambda x: x[0])
    ret = [ivs[0]]
    for iv in ivs[1:]:
        if iv[0] - ret[-1][1] > dist:
            ret.append(iv)
        else:
            ret[-1] = (ret[-1][0], max(iv[1], ret[-1][1]))
    return ret


class Pair:
    """utils/calignment.pyx AlignedPair (the parts the loader reads)."""

    def __init__(self, r1, r2=None):
        self.r1, self.r2 = r1, r2

    query_name = property(lambda s: s.r1.qname)
    is_unmapped = property(lambda s: s.r1.is_unmapped)
    ref_id = property(lambda s: s.r1.ref_id)

    @property
    def refblocks(self):                       # calignment.pyx:96-102
        if self.r2 is None:
            return merge_blocks(self.r1.get_blocks(), 1)
        return merge_blocks(self.r1.get_blocks() + self.r2.get_blocks(), 1)

    @property
    def alnlen(self):                          # :103-105
        return sum(b[1] - b[0] for b in self.refblocks)

    @property
    def alnscore(self):                        # :107-111
        if self.r2 is None:
            return self.r1.tags['AS']
        return self.r1.tags['AS'] + self.r2.tags['AS']


def _readkey(a):                               # alignment.py:34-58
    return (a.qname, a.is_read1, a.ref_id, a.pos, a.next_ref_id, a.next_pos, a.tlen)


def _matekey(a):                               # alignment.py:61-90
    return (a.qname, not a.is_read1, a.next_ref_id, a.next_pos, a.ref_id, a.pos, -a.tlen)


def pair_bundle(alns):                         # alignment.py:181-198
    cache = {}
    for a in alns:
        if not a.is_paired:
            yield Pair(a)
        else:
            mate = cache.pop(_matekey(a), None)
            if mate is not None:
                yield Pair(a, mate) if a.is_read1 else Pair(mate, a)
            else:
                cache[_readkey(a)] = a
    for a in cache.values():
        yield Pair(a)


def fetch_fragments_seq(alns):                 # alignment.py:171-214 (bundles of equal query name)
    i = 0
    while i < len(alns):
        j = i
        while j < len(alns) and alns[j].qname == alns[i].qname:
            j += 1
        b = alns[i:j]
        i = j
        if not b[0].is_paired:
            yield ('SU' if b[0].is_unmapped else 'SM'), [Pair(a) for a in b]
        elif b[0].is_proper_pair:
            yield 'PM', list(pair_bundle(b))
        elif len(b) == 2 and all(a.is_unmapped for a in b):
            yield 'PU', [Pair(b[0], b[1])]
        else:
            yield 'PX', [Pair(a) for a in b]


# ---------------------------------------------------------------------------------------------
# annotation (annotation/_intervaltree.py, annotation/utils.py)
# ---------------------------------------------------------------------------------------------
class Annotation:
    def __init__(self, gtf_file, attribute_name='locus', feature_type='exon'):
        self.loci = defaultdict(list)          # locus -> [(chrom, start, end)]
        for line in open(gtf_file):
            if line.startswith('#'):
                continue
            f = [x.strip() for x in line.split('\t')]
            if len(f) < 9 or f[2] != feature_type:
                continue
            attrs = dict(re.findall(r'(\w+)\s+"(.+?)";', f[8]))
            if attribute_name not in attrs:
                continue
            self.loci[attrs[attribute_name]].append((f[0], int(f[3]), int(f[4])))
        self.ivs = defaultdict(list)           # chrom -> [(begin, end, locus)]
        for locid, rows in self.loci.items():
            by_chrom = defaultdict(list)
            for chrom, s, e in rows:
                by_chrom[chrom].append((s, e + 1))                       # _intervaltree.py:54
            for chrom, iv in by_chrom.items():
                for b, e in self._merge_neighbors(iv):                   # :56-59
                    self.ivs[chrom].append((b, e, locid))

    @staticmethod
    def _merge_neighbors(ivs, distance=1):     # annotation/utils.py:17-46
        ivs = sorted(set(ivs))
        merged = [ivs[0]]
        for b, e in ivs[1:]:
            lb, le = merged[-1]
            if b - le <= distance:
                merged[-1] = (lb, max(le, e))
            else:
                merged.append((b, e))
        return merged

    def intersect_blocks(self, ref, blocks):   # _intervaltree.py:76-82
        res = Counter()
        for b_start, b_end in blocks:
            qb, qe = b_start, b_end + 1
            for b, e, loc in self.ivs.get(ref, ()):
                if b < qe and qb < e:          # IntervalTree.overlap: strict overlap of half-open intervals
                    res[loc] += max(0, min(e, qe) - max(b, qb))
        return res


# ---------------------------------------------------------------------------------------------
# the loader (utils/model.py)
# ---------------------------------------------------------------------------------------------
def load_alignment(bam_path, gtf_path, barcode_tag='CB', ignore_umi=False, umi_tag='UB', overlap_threshold=0.2,
                   no_feature_key='__no_feature'):
    """Restates ``Stellarscope.load_alignment`` (model.py:946-1146) for the unstranded threshold
    overlap mode.  Returns a dict with the uint16 CSR score matrix and the index maps."""
    import scipy.sparse as sp
    refs, alns = read_bam(bam_path)
    annot = Annotation(gtf_path)
    feat_index = {no_feature_key: 0}
    for locus in annot.loci.keys():
        feat_index.setdefault(locus, len(feat_index))                   # model.py:987-989

    def assign(pair):                                                   # model.py:519-537
        f = annot.intersect_blocks(refs[pair.ref_id], pair.refblocks)
        if not f:
            return no_feature_key
        fname, overlap = f.most_common()[0]
        return fname if overlap > pair.alnlen * overlap_threshold else no_feature_key

    read_index, read_bcode, read_umi = {}, {}, {}
    bcode_ridx = defaultdict(set)
    mappings = []
    min_as, max_as = np.iinfo(np.int32).max, np.iinfo(np.int32).min
    n_frag = 0
    for code, pairs in fetch_fragments_seq(alns):
        n_frag += 1
        if code in ('SU', 'PU'):
            continue
        tags = [p.r1.tags.get(barcode_tag) for p in pairs]              # get_tag_alignments, alignment.py:277-291
        bcode = next((t for t in tags if t is not None), None)
        umi = next((p.r1.tags.get(umi_tag) for p in pairs if p.r1.tags.get(umi_tag) is not None), None)
        if bcode is None:
            continue                                                    # "Missing CB", model.py:1099-1101
        if not ignore_umi and umi is None:
            continue
        mapped = [p for p in pairs if not p.is_unmapped]
        scores = [p.alnscore for p in mapped]
        feats = [assign(p) for p in mapped]
        min_as, max_as = min([min_as] + scores), max([max_as] + scores)  # statistics.py:230-231
        if not any(f != no_feature_key for f in feats):
            continue
        q = mapped[0].query_name
        row = read_index.setdefault(q, len(read_index))
        read_bcode.setdefault(q, bcode)
        bcode_ridx[bcode].add(row)
        if not ignore_umi:
            read_umi.setdefault(q, umi)
        byfeat = defaultdict(list)                                      # process_overlap_frag, model.py:51-84
        for p, f in zip(mapped, feats):
            byfeat[f].append(p)
        maps = []
        for f, fal in byfeat.items():
            fal.sort(key=lambda x: x.alnscore + x.alnlen, reverse=True)
            maps.append((q, f, fal[0].alnscore, fal[0].alnlen))
        maps.sort(key=lambda x: x[2], reverse=True)
        mappings.extend(maps)
    shape = (len(read_index), len(feat_index))
    dense = {}
    for q, f, alnscore, alnlen in mappings:                             # mapping_to_matrix, model.py:958-978
        key = (read_index[q], feat_index[f])
        dense[key] = max(dense.get(key, 0), (alnscore - min_as + 1) + alnlen)
    rows = np.fromiter((k[0] for k in dense), dtype=np.int64, count=len(dense))
    cols = np.fromiter((k[1] for k in dense), dtype=np.int64, count=len(dense))
    vals = np.fromiter(dense.values(), dtype=np.int64, count=len(dense))
    assert len(vals) == 0 or vals.max() < 65536
    m = sp.csr_matrix((vals.astype(np.uint16), (rows, cols)), shape=shape)
    m.sort_indices()
    return dict(matrix=m, read_index=read_index, feat_index=feat_index, read_bcode=read_bcode,
                bcode_ridx={k: sorted(v) for k, v in bcode_ridx.items()}, minAS=int(min_as), maxAS=int(max_as),
                n_fragments=n_frag)
