"""Decoder fuzz (dev container: needs oracle/_ref): corrupted frames through the kernel source on the host (fused + staged decoders) vs the compiled reference; prints every difference in result / error code.  usage: python scripts/fuzz_decoder_vs_ref.py <seed> <seconds>"""
import sys, random, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from tests.oracle_util import *
from zstd_jni_b200 import corpus
seed=int(sys.argv[1]); T=float(sys.argv[2])
rnd=random.Random(seed)
srcs=[corpus.chunk(i).tobytes() for i in range(8)]+[corpus.chunk(i)[:rnd.randint(1,60000)].tobytes() for i in range(8)]
frames=[]
for d in srcs:
    for lvl in (1,3,6,9):
        frames.append((ref_compress_flags(d,lvl,rnd.random()<0.3,rnd.random()<0.8,False),len(d),False))
    frames.append((ref_compress_flags(d,3,False,True,True),len(d),True))
multi=ref_stream_compress(b"".join(srcs[:3]),3)
frames.append((multi,3*131072,False))
t0=time.time(); n=0; bad=0; strict=0
while time.time()-t0<T:
    f,cap,ml=rnd.choice(frames)
    b=bytearray(f)
    k=rnd.choice([1,1,1,2,3,8])
    mode=rnd.random()
    if mode<0.7:
        for _ in range(k):
            i=rnd.randrange(len(b)); b[i]^=1<<rnd.randrange(8) if rnd.random()<0.5 else rnd.randrange(1,256)
    elif mode<0.85:
        b=b[:rnd.randrange(len(b))]
    else:
        i=rnd.randrange(len(b)); b[i:i+rnd.randint(1,4)]=bytes(rnd.randrange(256) for _ in range(rnd.randint(0,6)))
    b=bytes(b)
    capx=cap if rnd.random()<0.8 else rnd.randrange(cap+1)
    if ml:
        e=ref_decompress_magicless(b,capx); g=hostsim_decompress_magicless(b,capx)
    else:
        e=ref_decompress(b,capx); g=hostsim_decompress(b,capx)
        if rnd.random()<0.3:
            g2=staged_decompress(b,capx)
            if g2!=e:
                if g2==-20: strict+=1
                else: bad+=1; print("STAGED MISMATCH",seed,n,len(b), e if isinstance(e,int) else len(e), g2 if isinstance(g2,int) else len(g2)); open(f"/tmp/fuzz_bad_{seed}_{n}.bin","wb").write(b)
    n+=1
    if e!=g and g==-20: strict+=1      # the documented divergence: strict end-of-stream checks vs the reference's lenient fast Huffman loop
    elif e!=g:
        bad+=1; print("MISMATCH",seed,n,ml,len(b),capx, e if isinstance(e,int) else len(e), g if isinstance(g,int) else len(g)); open(f"/tmp/fuzz_bad_{seed}_{n}.bin","wb").write(b)
print("seed",seed,"cases",n,"stricter-than-reference (corruption_detected where the reference went on)",strict,"other differences",bad)
