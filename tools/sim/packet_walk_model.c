// CPU model of the packet walk (profiles/round2_summary.md section 3): how many nodes does ONE traversal shared by
// W rays visit on the real BVH and rays?  It decided the packet width, the 8 x 8 tiling and the static child order before the kernels
// were written, and it is where the "not built" estimates (4-wide collapse, dead pops) come from.
//
//   python tools/sim/dump_inputs.py /tmp/sim            (CPU oracle: sponza_q.bin = quality BVH, sponza_f.bin = LBVH, rays4k.bin)
//   gcc -O2 -o /tmp/sim/model tools/sim/packet_walk_model.c -lm
//   cd /tmp/sim && ./model sponza_q.bin rays4k.bin W L [TW]
//     W  rays per packet (64), L levels of the tree collapsed into one wide node (1 = the binary tree, 2 = 4-wide, 3 = 8-wide),
//     TW tile width in pixels of a 3840-wide image (W = strips of W consecutive rays, 8 = 8 x 8 tiles for W = 64)
//   prints, per packet: wide-node visits, child boxes tested, children some ray wants, pushes, leaves visited, deepest stack.
// Children are ordered statically per direction octant along the axis of largest centre distance (rr_internal.h node_order_bits);
// a packet is entered only when all its rays share an octant.  Same slab test / triangle test arithmetic as the oracle.
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <math.h>
#include <string.h>
typedef struct { float a0[3]; uint32_t c0; float a1[3]; uint32_t c1; float b0[3]; uint32_t parent; float b1[3]; uint32_t upd; } node;
typedef struct { float o[3]; float tmin; float d[3]; float tmax; } ray;
static float safe_inv(float d){ const float e=1e-5f; return 1.0f/(fabsf(d)>e?d:(d<0?-e:e)); }
static int slab(const float*mn,const float*mx,const float*inv,const float*ox,float tmax,float tmin,float*t0o){
  float t1=tmax,t0=tmin;
  for(int a=0;a<3;a++){ float f=fmaf(mx[a],inv[a],ox[a]), n=fmaf(mn[a],inv[a],ox[a]); t1=fminf(t1,fmaxf(f,n)); t0=fmaxf(t0,fminf(f,n)); }
  *t0o=t0; return t0<=t1; }
static int tri(const ray*r,const float*v0,const float*v1,const float*v2,float tmax,float*t){
  float e1[3],e2[3],s1[3],dd[3],s2[3];
  for(int i=0;i<3;i++){e1[i]=v1[i]-v0[i];e2[i]=v2[i]-v0[i];dd[i]=r->o[i]-v0[i];}
  s1[0]=r->d[1]*e2[2]-e2[1]*r->d[2]; s1[1]=r->d[2]*e2[0]-e2[2]*r->d[0]; s1[2]=r->d[0]*e2[1]-e2[0]*r->d[1];
  float den=(s1[0]*e1[0]+s1[1]*e1[1])+s1[2]*e1[2]; if(den==0) return 0; float inv=1.0f/den;
  float b1=((dd[0]*s1[0]+dd[1]*s1[1])+dd[2]*s1[2])*inv;
  s2[0]=dd[1]*e1[2]-e1[1]*dd[2]; s2[1]=dd[2]*e1[0]-e1[2]*dd[0]; s2[2]=dd[0]*e1[1]-e1[0]*dd[1];
  float b2=((r->d[0]*s2[0]+r->d[1]*s2[1])+r->d[2]*s2[2])*inv;
  float tt=((e2[0]*s2[0]+e2[1]*s2[1])+e2[2]*s2[2])*inv;
  if(b1<0||b1>1||b2<0||b1+b2>1||tt<r->tmin||tt>tmax) return 0; *t=tt; return 1; }
typedef struct { const float*mn,*mx; uint32_t ref; } child;
static node*N; static int oct;
// expand node i to 'levels' levels, in octant order (near first); returns count
static int expand(uint32_t i,int levels,child*out){
  node*n=&N[i]; 
  // order of the two children by the axis of largest centroid separation
  int ax=0; float best=-1; for(int a=0;a<3;a++){ float d=fabsf((n->b0[a]+n->b1[a])-(n->a0[a]+n->a1[a])); if(d>best){best=d;ax=a;} }
  int c0first = ((n->a0[ax]+n->a1[ax]) <= (n->b0[ax]+n->b1[ax])); if((oct>>ax)&1) c0first=!c0first;  // oct bit set = negative direction
  int cnt=0;
  for(int k=0;k<2;k++){ int which = (k==0)?(c0first?0:1):(c0first?1:0);
    uint32_t c = which?n->c1:n->c0; const float*mn=which?n->b0:n->a0,*mx=which?n->b1:n->a1;
    if(levels>1 && N[c].c0!=~0u) cnt+=expand(c,levels-1,out+cnt); else { out[cnt].mn=mn; out[cnt].mx=mx; out[cnt].ref=c; cnt++; } }
  return cnt; }
int main(int argc,char**argv){
  FILE*f=fopen(argv[1],"rb"); fseek(f,0,SEEK_END); long sz=ftell(f); fseek(f,0,SEEK_SET); N=malloc(sz); fread(N,1,sz,f); fclose(f);
  f=fopen(argv[2],"rb"); fseek(f,0,SEEK_END); long rs=ftell(f); fseek(f,0,SEEK_SET); ray*R=malloc(rs); fread(R,1,rs,f); fclose(f);
  long nr=rs/32; int W=atoi(argv[3]); int L=atoi(argv[4]); int TW=argc>5?atoi(argv[5]):W; int TH=W/TW; int IW=3840;
  double Ui=0,Ul=0,Ub=0,Upush=0,Uhitch=0; long packets=0, skipped=0; int maxsp=0;
  for(long p=0;p+W<=nr;p+=W*37){
    float inv[256][3],ox[256][3],cl[256]; long RI[256]; uint32_t cp[256]; int o0=-1, same=1;
    for(int l=0;l<W;l++){ long pk=p/W; long tpr=IW/TW; long ty=pk/tpr, tx=pk%tpr; long ridx=(ty*TH+l/TW)*IW+tx*TW+l%TW; if(ridx>=nr) ridx=nr-1; RI[l]=ridx; ray*r=&R[ridx]; int o=0; for(int a=0;a<3;a++){inv[l][a]=safe_inv(r->d[a]); ox[l][a]=-r->o[a]*inv[l][a]; if(r->d[a]<0) o|=1<<a;} cl[l]=r->tmax; cp[l]=~0u; if(l==0)o0=o; else if(o!=o0) same=0; }
    if(!same){skipped++;continue;}
    oct=o0;
    uint32_t st[1024]; int sp=0; st[sp++]=~0u; uint32_t a=0; long ni=0,nl=0,nb=0,npush=0,nh=0;
    while(a!=~0u){ node*n=&N[a];
      if(n->c0!=~0u){ ni++; child ch[16]; int cnt=expand(a,L,ch); nb+=cnt; int hit[16]; int nhit=0;
        for(int k=0;k<cnt;k++){ int any=0; for(int l=0;l<W;l++){ float t0; if(slab(ch[k].mn,ch[k].mx,inv[l],ox[l],cl[l],R[RI[l]].tmin,&t0)){any=1;break;} } hit[k]=any; nhit+=any; }
        nh+=nhit;
        // push hit children in reverse order, then pop
        for(int k=cnt-1;k>=0;k--) if(hit[k]){ st[sp++]=ch[k].ref; }
        if(nhit>1) npush+=nhit-1; if(sp>maxsp)maxsp=sp;
      }
      else { nl++; for(int l=0;l<W;l++){ float t; if(tri(&R[RI[l]],n->a0,n->a1,n->b0,cl[l],&t)){ if(t<cl[l]||(t==cl[l]&&n->c1<cp[l])){cl[l]=t;cp[l]=n->c1;} } } }
      a=st[--sp]; }
    Ui+=ni; Ul+=nl; Ub+=nb; Upush+=npush; Uhitch+=nh; packets++;
  }
  printf("L=%d W=%d packets %ld (skipped %ld): wide visits %.1f boxes %.1f hit-children %.1f pushes %.1f leaves %.1f maxsp %d\n",L,W,packets,skipped,Ui/packets,Ub/packets,Uhitch/packets,Upush/packets,Ul/packets,maxsp);
  return 0; }
