// tools/experiments/node_quant_probe.cpp -- host probe (test-only code path: compiles tests/host_emul/emul.cpp): node visits and
// leaf candidates per ray for exact fp32 / fp16 outward / 8-bit parent-local child boxes on primary and cosine-scattered rays.
// g++ -O2 -std=c++17 -fopenmp -ffp-contract=off -I/usr/local/cuda/include -Iraytracing-course_b200/csrc -Iinclude \
//     tools/experiments/node_quant_probe.cpp raytracing-course_b200/csrc/build/{scene_load,bvh_build}.o -o /tmp/probe && /tmp/probe scenes/practice5_dragon_100k.txt
#include "../../tests/host_emul/emul.cpp"
#include <cstdio>
#include <vector>
#include <cmath>
#include <random>
#include <functional>
static float h2f(uint16_t h){ __half x; memcpy(&x,&h,2); return __half2float(x);} 
struct Node { float lo[4][3], hi[4][3]; uint32_t ref[4]; };
static float half_dn(float v){ __half h=__float2half_rd(v); return __half2float(h);} 
static float half_up_(float v){ __half h=__float2half_ru(v); return __half2float(h);} 
int main(int argc,char**argv){
  EmuScene* e=(EmuScene*)emu_scene_load(argv[1]);
  const FlatScene& F=e->host.flat; DevScene& S=e->dev;
  size_t nn=F.inodes.size()/kIndexNodeF4;
  std::vector<Node> ex(nn);
  // decode refs
  for(size_t i=0;i<nn;i++){ const uint32_t* w=(const uint32_t*)&F.inodes[kIndexNodeF4*i]; for(int c=0;c<4;c++) ex[i].ref[c]=w[12+c]; }
  // exact boxes bottom-up (recursive)
  std::function<void(uint32_t,float*,float*)> boxof=[&](uint32_t ref,float*lo,float*hi){
    if(ref&IREF_LEAF){ uint32_t first=ref&0xFFFFFF; const f4&a=F.ubox[2*first],&b=F.ubox[2*first+1]; lo[0]=a.x;lo[1]=a.y;lo[2]=a.z;hi[0]=b.x;hi[1]=b.y;hi[2]=b.z; return;}
    Node&n=ex[ref]; for(int k=0;k<3;k++){lo[k]=1e30f;hi[k]=-1e30f;}
    for(int c=0;c<4;c++){ if(n.ref[c]==IREF_NONE){ for(int k=0;k<3;k++){n.lo[c][k]=6e4f;n.hi[c][k]=6e4f;} continue;} boxof(n.ref[c],n.lo[c],n.hi[c]); for(int k=0;k<3;k++){lo[k]=fminf(lo[k],n.lo[c][k]);hi[k]=fmaxf(hi[k],n.hi[c][k]);} }
  };
  float rl[3],rh[3]; boxof(S.iroot,rl,rh);
  printf("nodes %zu root box %g %g %g - %g %g %g\n",nn,rl[0],rl[1],rl[2],rh[0],rh[1],rh[2]);
  // quantised variants
  auto make=[&](int mode,int margin){ std::vector<Node> q=ex; for(size_t i=0;i<nn;i++){ Node&n=q[i];
      if(mode==1){ for(int c=0;c<4;c++) for(int k=0;k<3;k++){ n.lo[c][k]=half_dn(n.lo[c][k]); n.hi[c][k]=half_up_(n.hi[c][k]); } }
      if(mode==2){ float lo[3]={1e30f,1e30f,1e30f},hi[3]={-1e30f,-1e30f,-1e30f}; for(int c=0;c<4;c++) if(n.ref[c]!=IREF_NONE) for(int k=0;k<3;k++){lo[k]=fminf(lo[k],n.lo[c][k]);hi[k]=fmaxf(hi[k],n.hi[c][k]);}
        for(int k=0;k<3;k++){ float ext=hi[k]-lo[k]; int ee=(int)ceilf(log2f(fmaxf(ext,1e-30f)/(255.f-2*margin))); if(ee<-20)ee=-20; float sc=ldexpf(1.f,ee); float org=lo[k]-margin*sc;
          for(int c=0;c<4;c++) if(n.ref[c]!=IREF_NONE){ float a=floorf((n.lo[c][k]-org)/sc)-margin; float b=ceilf((n.hi[c][k]-org)/sc)+margin; if(a<0)a=0; if(b>255)b=255; n.lo[c][k]=org+a*sc; n.hi[c][k]=org+b*sc; } } }
    } return q; };
  // rays
  int W=S.width,H=S.height; std::vector<float> o,d; int stride=4;
  for(int y=0;y<H;y+=stride)for(int x=0;x<W;x+=stride){ vec3 ro,rd; camera_ray(S,x+0.5f,y+0.5f,ro,rd); o.insert(o.end(),{ro.x,ro.y,ro.z}); d.insert(d.end(),{rd.x,rd.y,rd.z}); }
  long n=o.size()/3; std::vector<float> so,sd; std::mt19937 rng(1); std::normal_distribution<float> N01;
  for(long i=0;i<n;i++){ uint32_t v=0,f=0; SceneHit h=scene_intersect<0>(S,mk3(o[3*i],o[3*i+1],o[3*i+2]),mk3(d[3*i],d[3*i+1],d[3*i+2]),&v,nullptr,&f); if(h.id<0)continue;
    vec3 p=mk3(o[3*i],o[3*i+1],o[3*i+2])+h.t*mk3(d[3*i],d[3*i+1],d[3*i+2]);
    for(int s=0;s<4;s++){ vec3 r=normalize(mk3(N01(rng),N01(rng),N01(rng))); vec3 dir=normalize(r+h.n); if(dot(dir,h.n)<=0) continue; vec3 q=p+1e-4f*dir; so.insert(so.end(),{q.x,q.y,q.z}); sd.insert(sd.end(),{dir.x,dir.y,dir.z}); } }
  printf("primary %ld secondary %ld\n",n,(long)so.size()/3);
  auto run=[&](const std::vector<Node>&q,const std::vector<float>&O,const std::vector<float>&D,const char*label){ long nr=O.size()/3; double visits=0,cands=0,pass=0,entered=0;
    #pragma omp parallel for reduction(+:visits,cands,pass,entered) schedule(dynamic,256)
    for(long i=0;i<nr;i++){ vec3 ro=mk3(O[3*i],O[3*i+1],O[3*i+2]),rd=mk3(D[3*i],D[3*i+1],D[3*i+2]); vec3 inv=mk3(1.f/rd.x,1.f/rd.y,1.f/rd.z),oi=ro*inv; uint32_t st[256];int sp=0; st[sp++]=S.iroot; long v=0;
      bool first=true; bool ent=false;
      while(sp){ uint32_t ref=st[--sp]; if(ref&IREF_LEAF){ cands++; uint32_t fi=ref&0xFFFFFF; const f4&a=F.ubox[2*fi],&b=F.ubox[2*fi+1]; bool hit;float tc; slab(a.x,a.y,a.z,b.x,b.y,b.z,inv,oi,ref,hit,tc); if(hit)pass++; continue;}
        const Node&nd=q[ref]; v++; bool any=false; for(int c=0;c<4;c++){ bool hit;float tc; slab(nd.lo[c][0],nd.lo[c][1],nd.lo[c][2],nd.hi[c][0],nd.hi[c][1],nd.hi[c][2],inv,oi,nd.ref[c],hit,tc); if(hit){st[sp++]=nd.ref[c];any=true;} }
        if(first){first=false; ent=any;} }
      if(ent){entered++; visits+=v;} else visits+=0; }
    printf("%-28s rays %ld entered %.0f visits/entered %.2f cands/entered %.2f exactpass/entered %.2f\n",label,nr,entered,visits/entered,cands/entered,pass/entered); };
  for(int pass=0;pass<2;pass++){ const auto&O=pass?so:o; const auto&D=pass?sd:d; printf("-- %s\n",pass?"secondary":"primary");
    run(ex,O,D,"fp32 exact"); run(make(1,0),O,D,"fp16 outward (current)"); run(make(2,0),O,D,"8-bit local margin 0"); run(make(2,1),O,D,"8-bit local margin 1"); }
}
