import sys, numpy as np
import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,R); sys.path.insert(0,os.path.join(R,'tests'))
from genomeassembler_dev_b200 import breakscore as B, tables, synth
import parity_cases as P
from oracle import loader as O
kmers=tables.all_kmer_strings(); prob=tables.normalised(tables.load_raw())
sc=B.BreakageScorer(0,'/tmp/asan/libbreakscore_emul.so')
import conftest
n=0
for mode in (0,B.PLACE_TILE,B.PLACE_SCAN):
    for params in P.SMALL[:3]:
        P.check_segment(sc,O,kmers,prob,P.make(*params,mut=0.5),flags=P.FULL|mode|B.WANT_LEV); n+=1
    for name,contigs,reads,truth,kmer in P.edge_inputs():
        P.check_segment(sc,O,kmers,prob,synth.Segment(truth,None,contigs),kmer=kmer,reads=reads,flags=P.FULL|mode|B.WANT_LEV); n+=1
rowid=np.arange(1,len(prob)+1,dtype=np.float64)
P.check_segment(sc,O,kmers,rowid,P.make(*P.SMALL[0]),flags=P.FULL); n+=1
b=synth.make_batch(5,seed=3,length=1500,read_len=30,coverage=6,contigs_lo=1,contigs_hi=4)
sc.set_table(kmers,prob)
sc.score_batch(b.read_chars,None,b.read_len,b.contig_chars,b.contig_off,b.truth_chars,b.truth_off,b.seg_read_start,b.seg_contig_start,flags=B.DEFAULT_FLAGS|B.WANT_HIST|B.WANT_POS|B.WANT_LEV); n+=1
from genomeassembler_dev_b200 import tables as T
sc.set_table(kmers,prob); sc.set_second_table(T.uniform(len(prob)))
seg=P.make(43, 2500, 50, 10, 5, 1)
sc.score(seg.contigs, seg.read_list, seg.truth, flags=B.DEFAULT_FLAGS|B.WANT_SECOND_TABLE|B.WANT_LEV|B.WANT_HIST); n+=1
reads,srs=sc.simulate_reads([seg.truth, b'ACGT', b''], 40, 8, seed=3); n+=1
print(B.assemble_contigs([b'ACGTACGTAAGGCCTT', b'GGCCTTACGTTTTTTTTT', b'TTTTTTTTTGGA'], 7, 5, n_shuffles=50, lib_path='/tmp/asan/libbreakscore_emul.so')); n+=1
# hashed placement scratch, smallest table first (launch repeated with larger tables)
os.environ['BS_PLACE_SCRATCH_MB']='0'; os.environ['BS_PLACE_HASH_CAP']='1'
for params in P.SMALL[:4]:
    P.check_segment(sc,O,kmers,prob,P.make(*params,mut=0.5),flags=P.FULL); n+=1
for name,contigs,reads,truth,kmer in P.edge_inputs():
    P.check_segment(sc,O,kmers,prob,synth.Segment(truth,None,contigs),kmer=kmer,reads=reads,flags=P.FULL); n+=1
del os.environ['BS_PLACE_SCRATCH_MB'], os.environ['BS_PLACE_HASH_CAP']
# one segment over several contexts (bs_score_multi: host-side sharding and scatter), every output kind, odd contig sets
others=[B.BreakageScorer(0,'/tmp/asan/libbreakscore_emul.so') for _ in range(2)]
for o in others: o.set_table(kmers,prob); o.set_second_table(T.uniform(len(prob)))
sc.set_table(kmers,prob); sc.set_second_table(T.uniform(len(prob)))
fl=B.DEFAULT_FLAGS|B.WANT_HIST|B.WANT_POS|B.WANT_LEV|B.WANT_SECOND_TABLE
for ctgs in (list(seg.contigs)+[b'',seg.contigs[0][:3]], [seg.contigs[0]], [b'',b''], list(seg.contigs)*2):
    a=sc.score(ctgs, seg.read_list, seg.truth, flags=fl); g=sc.score(ctgs, seg.read_list, seg.truth, flags=fl, group=others)
    assert all(np.array_equal(a[k],g[k],equal_nan=True) for k in a if k not in ('sequence','path_prob_dist','path_prob_dist2')); n+=1
sc.score(list(seg.contigs), [], seg.truth, flags=fl, group=others); n+=1   # no reads at all
for o in others: o.close()
# prefix-bitmap scan of the contig-in-truth search; interrupt poll between chunks
os.environ['BS_STARTPOS_BIG']='1'
for params in P.SMALL[:3]:
    P.check_segment(sc,O,kmers,prob,P.make(*params,mut=0.5),flags=P.FULL|B.WANT_LEV); n+=1
for name,contigs,reads,truth,kmer in P.edge_inputs():
    P.check_segment(sc,O,kmers,prob,synth.Segment(truth,None,contigs),kmer=kmer,reads=reads,flags=P.FULL|B.WANT_LEV); n+=1
del os.environ['BS_STARTPOS_BIG']
os.environ['BS_CHUNK_KB']='12'
with B.BreakageScorer(0,'/tmp/asan/libbreakscore_emul.so') as sc2:
    sc2.set_table(kmers,prob); calls=[]
    sc2.set_poll(lambda: calls.append(1) or len(calls)>=2)
    try:
        sc2.score_batch(b.read_chars,None,b.read_len,b.contig_chars,b.contig_off,b.truth_chars,b.truth_off,b.seg_read_start,b.seg_contig_start); raise SystemExit('not interrupted')
    except B.BreakscoreError as e: assert e.code==B.ERR_INTERRUPTED
    sc2.set_poll(None)
    sc2.score_batch(b.read_chars,None,b.read_len,b.contig_chars,b.contig_off,b.truth_chars,b.truth_off,b.seg_read_start,b.seg_contig_start); n+=1
del os.environ['BS_CHUNK_KB']
# scaffold sets scored from their parts (bs_score_scaffolds): every mode of the compositional path, hand-built and random sets
import scaffold_cases as SC
L='/tmp/asan/libbreakscore_emul.so'
modes=[{}, {'BS_COMPOSE_SCORE':'0'}, {'BS_COMPOSE_ROWS':'global'}, {'BS_COMPOSE_HASH_SLOTS':'64'}, {'BS_COMPOSE_TEXT':'1'}, {'BS_COMPOSE_JUNCTIONS':'0'}]
for env in modes:
    os.environ.update(env)
    for flags in (SC.FULL, SC.FULL & ~B.WANT_PROB_DIST, SC.FULL | B.WANT_LEV):
        for name,base,chains,reads,truth,kmer in SC.hand_sets():
            SC.check_scaffolds(sc,O,kmers,prob,truth,reads,SC.hand_scaffold_set(base,chains,L),kmer=kmer,flags=flags); n+=1
        for kw in (dict(seed=61,length=1500,read_len=40,coverage=8,n_base=6,n_scaffolds=12,overlap=9), dict(seed=63,length=1200,read_len=33,coverage=8,n_base=8,n_scaffolds=10,overlap=15,ragged=True)):
            truth,reads,sset=SC.make_set(lib_path=L,**kw)
            SC.check_scaffolds(sc,O,kmers,prob,truth,reads,sset,flags=flags); n+=1
    for k in env: del os.environ[k]
truth,reads,sset=SC.make_set(65,length=1200,read_len=50,coverage=6,n_base=5,n_scaffolds=8,overlap=11,lib_path=L)
SC.check_scaffolds(sc,O,kmers,prob,truth,reads,sset,second=T.uniform(len(prob))); n+=1
strings,aset=B.assemble_scaffolds([b'ACGTACGTAAGGCCTT', b'GGCCTTACGTTTTTTTTT', b'TTTTTTTTTGGA'], 7, 5, n_shuffles=50, lib_path=L)
assert aset.texts()==strings; n+=1
print('asan run ok', n, 'cases')
