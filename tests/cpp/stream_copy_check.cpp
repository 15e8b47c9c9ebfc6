
#include <vector>
#include <cstdio>
#include <cstring>
#include <cstdint>
#include <cstdlib>
namespace bk { void stream_copy(void*, const void*, size_t); }
int main(){ std::vector<uint8_t> a(1<<22), b((1<<22)+64); for(size_t i=0;i<a.size();i++) a[i]=(uint8_t)(i*131u>>3);
 size_t sizes[]={0,1,31,32,33,4095,4096,4097,100000,(1<<22)-77}; int bad=0;
 for(size_t off=0; off<40; off+=13) for(size_t so=0; so<9; so+=4) for(size_t n: sizes){ if(so+n>a.size()) continue; memset(b.data(),0xEE,b.size()); bk::stream_copy(b.data()+off,a.data()+so,n);
   if(memcmp(b.data()+off,a.data()+so,n)) bad++; if(off && b[off-1]!=0xEE) bad++; if(b[off+n]!=0xEE) bad++; }
 printf("bad=%d\n",bad); return bad; }
