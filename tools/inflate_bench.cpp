// csrc/dmx_inflate.h against zlib on every BGZF block of a file: same bytes, time of each.
// build: g++ -O2 -o tools/inflate_bench tools/inflate_bench.cpp -lz ; usage: tools/inflate_bench FILE.bam
#include <zlib.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <chrono>
#include "../demuxlet_b200/csrc/dmx_inflate.h"
static double now(){return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();}
int main(int argc,char**argv){
  FILE*f=fopen(argv[1],"rb"); fseek(f,0,SEEK_END); long n=ftell(f); fseek(f,0,SEEK_SET); std::vector<uint8_t> c(n+16); if(fread(c.data(),1,n,f)!=(size_t)n) return 1; fclose(f);
  std::vector<uint8_t> o1(70000), o2(70000); size_t total=0, nblk=0, nfail=0, nmis=0; double t_z=0,t_f=0; z_stream zs; memset(&zs,0,sizeof zs); inflateInit2(&zs,-15);
  dmx_io::FastInflate* fi = new dmx_io::FastInflate();
  for(size_t off=0; off+28<=(size_t)n;){ const uint8_t*b=c.data()+off; size_t bs=(b[16]|(b[17]<<8))+1; uint32_t isz; memcpy(&isz,b+bs-4,4);
    double t0=now(); inflateReset(&zs); zs.next_in=(Bytef*)b+18; zs.avail_in=bs-26; zs.next_out=o1.data(); zs.avail_out=isz; int rc=inflate(&zs,Z_FINISH); double t1=now();
    memset(o2.data(),0xEE,isz+16);
    bool ok = fi->run(b+18, bs-26, o2.data(), isz); double t2=now();
    t_z+=t1-t0; t_f+=t2-t1; total+=isz; ++nblk;
    if(!ok) { ++nfail; if (nfail==1) { size_t k=0; while(k<isz && o1[k]==o2[k]) ++k; printf("first failure: block %zu isz %u first diff at %zu\n", nblk, isz, k);} } else if (memcmp(o1.data(),o2.data(),isz)) ++nmis;
    if (o2[isz]!=0xEE) { printf("OVERRUN at block %zu\n", nblk); return 2; }
    (void)rc; off+=bs; }
  printf("%s: %zu blocks %.1f MB: zlib %.3f s (%.0f MB/s), fast %.3f s (%.0f MB/s), fast failed %zu, mismatched %zu\n", argv[1], nblk, total/1e6, t_z, total/1e6/t_z, t_f, total/1e6/t_f, nfail, nmis);
}
