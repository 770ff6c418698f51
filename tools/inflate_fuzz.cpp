// csrc/dmx_inflate.h on damaged BGZF blocks (bit flips, wrong ISIZE, truncated input) in exact-size heap buffers: to be
// run under AddressSanitizer -- the decoder may accept a damaged block (the CRC-32 check of the caller is what rejects it)
// but must never read or write out of bounds.
// build: g++ -O1 -g -fsanitize=address,undefined -o tools/inflate_fuzz tools/inflate_fuzz.cpp -lz ; usage: ASAN_OPTIONS=detect_leaks=0 tools/inflate_fuzz FILE.bam NBLOCKS
#include <zlib.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../demuxlet_b200/csrc/dmx_inflate.h"
int main(int argc,char**argv){
  FILE*f=fopen(argv[1],"rb"); fseek(f,0,SEEK_END); long n=ftell(f); fseek(f,0,SEEK_SET); std::vector<uint8_t> c(n); if(fread(c.data(),1,n,f)!=(size_t)n) return 1; fclose(f);
  dmx_io::FastInflate* fi = new dmx_io::FastInflate();
  srand(7); size_t nblk=0, accepted=0, accepted_wrong=0, tries=0;
  for(size_t off=0; off+28<=(size_t)n && nblk < (size_t)atoi(argv[2]);){ const uint8_t*b=c.data()+off; size_t bs=(b[16]|(b[17]<<8))+1; uint32_t isz; memcpy(&isz,b+bs-4,4); ++nblk;
    const size_t in_n = bs-26;
    // reference output
    std::vector<uint8_t> ref(isz ? isz : 1); z_stream zs; memset(&zs,0,sizeof zs); inflateInit2(&zs,-15); zs.next_in=(Bytef*)b+18; zs.avail_in=in_n; zs.next_out=ref.data(); zs.avail_out=isz; inflate(&zs,Z_FINISH); inflateEnd(&zs);
    for(int t=0;t<40;t++){
      uint8_t* in = (uint8_t*)malloc(in_n+8); memcpy(in,b+18,in_n+8);       // exact sizes: ASan sees any overrun
      int nflip = 1 + rand()%3; if (t==0) nflip=0;
      for(int k=0;k<nflip;k++) in[rand()%(in_n?in_n:1)] ^= (uint8_t)(1<<(rand()%8));
      size_t out_n = isz; if (t%7==3 && isz>10) out_n = isz - 1 - rand()%10; if (t%7==5) out_n = isz + 1 + rand()%10;   // wrong ISIZE too
      size_t use_n = in_n; if (t%11==4 && in_n>10) use_n = in_n - 1 - rand()%8;                                        // truncated input
      uint8_t* out = (uint8_t*)malloc(out_n?out_n:1);
      bool ok = fi->run(in, use_n, out, out_n); ++tries;
      if (ok) { ++accepted; if (out_n!=isz || memcmp(out,ref.data(),isz)) ++accepted_wrong; }
      free(in); free(out);
    }
    off+=bs; }
  printf("%zu blocks, %zu runs: accepted %zu, of those different from zlib's output %zu (left to the CRC check)\n", nblk, tries, accepted, accepted_wrong);
}
