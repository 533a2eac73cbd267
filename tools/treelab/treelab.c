/* treelab.c -- CPU experiment bench for traversal-structure design (NOT product code, NOT the oracle):
 * counts node visits / box tests / leaf tests per ray for candidate trees over a recorded ray set.
 * Primitive tests go through the oracle (orc_hit_object) so culling matches the product's semantics. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../oracle/rrt_oracle.h"

typedef struct { float lo[3], hi[3]; } box_t;
static inline void box_merge(box_t *a, const box_t *b){for(int k=0;k<3;k++){if(b->lo[k]<a->lo[k])a->lo[k]=b->lo[k];if(b->hi[k]>a->hi[k])a->hi[k]=b->hi[k];}}
static inline double box_area(const box_t *b){double x=b->hi[0]-b->lo[0],y=b->hi[1]-b->lo[1],z=b->hi[2]-b->lo[2];return 2*(x*y+y*z+z*x);}
static inline box_t box_empty(void){box_t b;for(int k=0;k<3;k++){b.lo[k]=INFINITY;b.hi[k]=-INFINITY;}return b;}

/* binary tree: child >=0 internal node index, <0 leaf ~prim */
typedef struct { int n; int *left, *right; box_t *box; /* [n-1] */ const box_t *pbox; int root; } btree;

static box_t refit(btree *t, int c){ if(c<0) return t->pbox[~c]; box_t a=refit(t,t->left[c]), b=refit(t,t->right[c]); box_merge(&a,&b); t->box[c]=a; return a; }

/* ---- SAH builder (full sweep for small ranges, binned for large) ---- */
static const box_t *g_pb; static int g_axis;
static int cmp_centroid(const void *a,const void *b){int i=*(const int*)a,j=*(const int*)b;float ci=g_pb[i].lo[g_axis]+g_pb[i].hi[g_axis],cj=g_pb[j].lo[g_axis]+g_pb[j].hi[g_axis];return ci<cj?-1:(ci>cj?1:(i<j?-1:(i>j)));}
int g_sweep_limit=4096, g_bins=32;
void lab_set_sah(int sweep_limit,int bins){g_sweep_limit=sweep_limit;g_bins=bins;}
static int sah_rec(btree *t, int *ids, int n, int *next){
  if(n==1) return ~ids[0];
  int node=(*next)++;
  int best_axis=-1,best_split=-1; double best=INFINITY;
  if(n<=g_sweep_limit){
    double *ra=malloc(sizeof(double)*n);
    for(int ax=0;ax<3;ax++){
      g_axis=ax; g_pb=t->pbox; qsort(ids,n,sizeof(int),cmp_centroid);
      box_t b=box_empty(); for(int i=n-1;i>0;i--){box_merge(&b,&t->pbox[ids[i]]);ra[i]=box_area(&b);}
      b=box_empty(); for(int i=0;i<n-1;i++){box_merge(&b,&t->pbox[ids[i]]);double c=box_area(&b)*(i+1)+ra[i+1]*(n-1-i); if(c<best){best=c;best_axis=ax;best_split=i+1;}}
    }
    free(ra);
    g_axis=best_axis; g_pb=t->pbox; qsort(ids,n,sizeof(int),cmp_centroid);
  } else {
    enum{NBMAX=64}; const int NB=g_bins;
    box_t cb=box_empty(); for(int i=0;i<n;i++){box_t c; for(int k=0;k<3;k++){c.lo[k]=c.hi[k]=0.5f*(t->pbox[ids[i]].lo[k]+t->pbox[ids[i]].hi[k]);} box_merge(&cb,&c);}
    int bb=-1;
    for(int ax=0;ax<3;ax++){
      float lo=cb.lo[ax],ext=cb.hi[ax]-cb.lo[ax]; if(!(ext>0)) continue;
      box_t bins[NBMAX]; int cnt[NBMAX]; for(int b=0;b<NB;b++){bins[b]=box_empty();cnt[b]=0;}
      for(int i=0;i<n;i++){float c=0.5f*(t->pbox[ids[i]].lo[ax]+t->pbox[ids[i]].hi[ax]);int b=(int)((c-lo)/ext*NB);if(b>=NB)b=NB-1;if(b<0)b=0;cnt[b]++;box_merge(&bins[b],&t->pbox[ids[i]]);}
      double ra[NBMAX]; int rc[NBMAX]; box_t b=box_empty(); int c=0; for(int k=NB-1;k>0;k--){box_merge(&b,&bins[k]);c+=cnt[k];ra[k]=box_area(&b);rc[k]=c;}
      b=box_empty(); c=0; for(int k=0;k<NB-1;k++){box_merge(&b,&bins[k]);c+=cnt[k]; if(c==0||rc[k+1]==0)continue; double cost=box_area(&b)*c+ra[k+1]*rc[k+1]; if(cost<best){best=cost;best_axis=ax;bb=k;}}
    }
    if(best_axis<0){ best_split=n/2; }
    else { float lo=cb.lo[best_axis],ext=cb.hi[best_axis]-cb.lo[best_axis]; int i=0,j=n-1; while(i<=j){float c=0.5f*(t->pbox[ids[i]].lo[best_axis]+t->pbox[ids[i]].hi[best_axis]);int b=(int)((c-lo)/ext*NB);if(b>=NB)b=NB-1;if(b<0)b=0; if(b<=bb)i++; else {int tmp=ids[i];ids[i]=ids[j];ids[j]=tmp;j--;}} best_split=i; if(best_split<=0||best_split>=n)best_split=n/2; }
  }
  int l=sah_rec(t,ids,best_split,next); int r=sah_rec(t,ids+best_split,n-best_split,next);
  t->left[node]=l; t->right[node]=r; return node;
}

btree *lab_tree_new(int n, const float *pbox6){ btree *t=calloc(1,sizeof(btree)); t->n=n; t->left=malloc(sizeof(int)*n); t->right=malloc(sizeof(int)*n); t->box=malloc(sizeof(box_t)*n); t->pbox=(const box_t*)pbox6; t->root=0; return t; }
btree *lab_build_sah(int n, const float *pbox6){ btree *t=lab_tree_new(n,pbox6); int *ids=malloc(sizeof(int)*n); for(int i=0;i<n;i++)ids[i]=i; int next=0; t->root=sah_rec(t,ids,n,&next); free(ids); refit(t,t->root); return t; }
/* from explicit arrays (canonical LBVH): leaf refs ~k index sorted position -> perm[k] = object id */
btree *lab_from_arrays(int n, const float *pbox6, const int *left, const int *right, const uint32_t *perm){ btree *t=lab_tree_new(n,pbox6); for(int i=0;i<n-1;i++){int l=left[i],r=right[i]; t->left[i]=l>=0?l:~(int)perm[~l]; t->right[i]=r>=0?r:~(int)perm[~r];} refit(t,0); return t; }
/* PLOC (Meister & Bittner 2018): clusters in Morton order; each round every cluster finds its nearest neighbour
 * (smallest merged surface area) within +-radius positions; mutual nearest neighbours merge; compact; repeat. */
btree *lab_build_ploc(int n, const float *pbox6, const uint32_t *perm, int radius){
  btree *t=lab_tree_new(n,pbox6); if(n==1){t->root=~0;return t;}
  int *cl=malloc(sizeof(int)*n), *nn=malloc(sizeof(int)*n), *out=malloc(sizeof(int)*n); box_t *cb=malloc(sizeof(box_t)*n), *ob=malloc(sizeof(box_t)*n);
  for(int i=0;i<n;i++){cl[i]=~(int)perm[i]; cb[i]=t->pbox[perm[i]];}
  int m=n, next=n-2; /* allocate internal nodes from the back so that the root ends up at index 0 */
  while(m>1){
    for(int i=0;i<m;i++){ double best=INFINITY; int bj=-1; int lo=i-radius<0?0:i-radius, hi=i+radius>=m?m-1:i+radius;
      for(int j=lo;j<=hi;j++){ if(j==i)continue; box_t b=cb[i]; box_merge(&b,&cb[j]); double a=box_area(&b); if(a<best){best=a;bj=j;} } nn[i]=bj; }
    int k=0;
    for(int i=0;i<m;i++){ int j=nn[i]; if(nn[j]==i){ if(i<j){ int node=next--; t->left[node]=cl[i]; t->right[node]=cl[j]; box_t b=cb[i]; box_merge(&b,&cb[j]); out[k]=node; ob[k]=b; k++; } } else { out[k]=cl[i]; ob[k]=cb[i]; k++; } }
    memcpy(cl,out,sizeof(int)*k); memcpy(cb,ob,sizeof(box_t)*k); m=k;
  }
  t->root=cl[0]; free(cl);free(nn);free(out);free(cb);free(ob); refit(t,t->root); return t; }
/* LBVH on top, full SAH below: every maximal LBVH subtree with <= max_leaves leaves is rebuilt by the SAH builder
 * (what one thread block per subtree would do on the GPU) */
static int count_leaves(const btree *t,int c){ return c<0?1:count_leaves(t,t->left[c])+count_leaves(t,t->right[c]); }
static void collect(const btree *t,int c,int *ids,int *n){ if(c<0){ids[(*n)++]=~c;return;} collect(t,t->left[c],ids,n); collect(t,t->right[c],ids,n); }
static int hybrid_rec(const btree *src,btree *dst,int c,int max_leaves,int *next){
  if(c<0) return c;
  int nl=count_leaves(src,c);
  if(nl<=max_leaves){ int *ids=malloc(sizeof(int)*nl); int k=0; collect(src,c,ids,&k); int r=sah_rec(dst,ids,nl,next); free(ids); return r; }
  int node=(*next)++; int l=hybrid_rec(src,dst,src->left[c],max_leaves,next); int r=hybrid_rec(src,dst,src->right[c],max_leaves,next); dst->left[node]=l; dst->right[node]=r; return node; }
btree *lab_build_hybrid(const btree *src,int max_leaves){ btree *t=lab_tree_new(src->n,(const float*)src->pbox); int next=0; t->root=hybrid_rec(src,t,src->root,max_leaves,&next); refit(t,t->root); return t; }
/* two levels: the subtrees of <= max_leaves leaves as above, then every maximal upper subtree with <= max_items of those
 * subtree roots (or single leaves) below it is rebuilt by the SAH builder over the ROOTS' boxes (one more thread block per
 * upper subtree on the GPU); only the few nodes above those keep the Morton topology */
static int count_items(const btree *t,int c,int max_leaves){ if(c<0) return 1; if(count_leaves(t,c)<=max_leaves) return 1; return count_items(t,t->left[c],max_leaves)+count_items(t,t->right[c],max_leaves); }
static void collect_items(const btree *src,btree *dst,int c,int max_leaves,int *next,int *refs,int *n){
  if(c<0){refs[(*n)++]=c;return;}
  int nl=count_leaves(src,c);
  if(nl<=max_leaves){ int *ids=malloc(sizeof(int)*nl); int k=0; collect(src,c,ids,&k); refs[(*n)++]=sah_rec(dst,ids,nl,next); free(ids); return; }
  collect_items(src,dst,src->left[c],max_leaves,next,refs,n); collect_items(src,dst,src->right[c],max_leaves,next,refs,n); }
static int graft(const btree *tmp,btree *dst,int c,const int *refs,int *next){ if(c<0) return refs[~c]; int node=(*next)++; int l=graft(tmp,dst,tmp->left[c],refs,next); int r=graft(tmp,dst,tmp->right[c],refs,next); dst->left[node]=l; dst->right[node]=r; return node; }
static int hybrid2_rec(const btree *src,btree *dst,int c,int max_leaves,int max_items,int *next){
  if(c<0) return c;
  int nl=count_leaves(src,c);
  if(nl<=max_leaves){ int *ids=malloc(sizeof(int)*nl); int k=0; collect(src,c,ids,&k); int r=sah_rec(dst,ids,nl,next); free(ids); return r; }
  int ni=count_items(src,c,max_leaves);
  if(ni<=max_items){
    int *refs=malloc(sizeof(int)*ni); int k=0; collect_items(src,dst,c,max_leaves,next,refs,&k);
    box_t *ib=malloc(sizeof(box_t)*ni); for(int i=0;i<ni;i++) ib[i]=refit(dst,refs[i]);
    btree *tmp=lab_tree_new(ni,(const float*)ib); int *ids=malloc(sizeof(int)*ni); for(int i=0;i<ni;i++)ids[i]=i; int tn=0; int troot=sah_rec(tmp,ids,ni,&tn);
    int r=graft(tmp,dst,troot,refs,next); free(ids); free(ib); free(refs); return r; }
  int node=(*next)++; int l=hybrid2_rec(src,dst,src->left[c],max_leaves,max_items,next); int r=hybrid2_rec(src,dst,src->right[c],max_leaves,max_items,next); dst->left[node]=l; dst->right[node]=r; return node; }
btree *lab_build_hybrid2(const btree *src,int max_leaves,int max_items){ btree *t=lab_tree_new(src->n,(const float*)src->pbox); int next=0; t->root=hybrid2_rec(src,t,src->root,max_leaves,max_items,&next); refit(t,t->root); return t; }
double lab_sah_cost(const btree *t){ double ra=box_area(&t->box[t->root]),c=0; for(int i=0;i<t->n-1;i++){ c+=box_area(&t->box[i])/ra*2.0; } return c; }

/* ---- ray/box ---- */
typedef struct { float o[3], d[3], inv[3], tm; } ray_t;
static inline int slab(const box_t *b,const ray_t *r,float tmin,float tmax,float *tent){ float t0=tmin,t1=tmax; for(int k=0;k<3;k++){float a=(b->lo[k]-r->o[k])*r->inv[k],c=(b->hi[k]-r->o[k])*r->inv[k]; if(a>c){float s=a;a=c;c=s;} if(a>t0)t0=a; if(c<t1)t1=c;} *tent=t0; return t0<=t1; }

typedef struct { uint64_t rays, visits, box_tests, leaf_tests, hits, pushes, maxstack; } lab_counters;

static inline int leaf_hit(const orc_scene *s,int prim,const float *ray7,float tmin,float *best_t,int *best_id,lab_counters *c){ float t; c->leaf_tests++; if(orc_hit_object(s,prim,ray7,tmin,*best_t,&t,NULL)){ if(t<*best_t||*best_id<0){*best_t=t;*best_id=prim;} return 1;} return 0; }

void lab_trace_binary(const btree *t,const orc_scene *s,const float *rays7,int m,float tmin,lab_counters *c,int32_t *ids){
  for(int i=0;i<m;i++){ const float *r7=rays7+7*i; ray_t r; for(int k=0;k<3;k++){r.o[k]=r7[k];r.d[k]=r7[3+k];r.inv[k]=1.0f/r.d[k];} r.tm=r7[6];
    float best=INFINITY; int bid=-1; int stack[128],sp=0; int cur=t->root; c->rays++;
    if(t->n==1){ leaf_hit(s,0,r7,tmin,&best,&bid,c); cur=INT32_MIN; }
    while(cur!=INT32_MIN){
      if(cur>=0){ c->visits++; c->box_tests+=2; int l=t->left[cur],rr=t->right[cur]; float tl,tr; const box_t *bl=l>=0?&t->box[l]:&t->pbox[~l],*br=rr>=0?&t->box[rr]:&t->pbox[~rr];
        int hl=slab(bl,&r,tmin,best,&tl),hr=slab(br,&r,tmin,best,&tr);
        if(hl&&hr){ if(tl<=tr){stack[sp++]=rr;cur=l;}else{stack[sp++]=l;cur=rr;} c->pushes++; if((uint64_t)sp>c->maxstack)c->maxstack=sp; }
        else if(hl)cur=l; else if(hr)cur=rr; else cur=sp?stack[--sp]:INT32_MIN; }
      else { leaf_hit(s,~cur,r7,tmin,&best,&bid,c); cur=sp?stack[--sp]:INT32_MIN; }
    }
    if(bid>=0)c->hits++; if(ids)ids[i]=bid; }
}

/* ---- wide tree collapsed from a binary tree ---- */
typedef struct { int k; int n_nodes; int *axis; int *child; /* [n_nodes*k] >=0 node, <0 leaf ~prim, INT32_MAX empty */ box_t *cbox; } wtree;
#define W_EMPTY 0x7fffffff
static int collapse_rec(const btree *t,wtree *w,int bnode){
  int me=w->n_nodes++; int k=w->k; int ch[16]; int nc=0; ch[nc++]=t->left[bnode]; ch[nc++]=t->right[bnode];
  while(nc<k){ int bi=-1; double ba=-1; for(int i=0;i<nc;i++) if(ch[i]>=0){double a=box_area(&t->box[ch[i]]); if(a>ba){ba=a;bi=i;}} if(bi<0)break; int b=ch[bi]; ch[bi]=t->left[b]; ch[nc++]=t->right[b]; }
  for(int i=0;i<k;i++){ if(i<nc){ w->cbox[me*k+i]= ch[i]>=0? t->box[ch[i]] : t->pbox[~ch[i]]; } else w->child[me*k+i]=W_EMPTY; }
  for(int i=0;i<nc;i++){ w->child[me*k+i]= ch[i]>=0? collapse_rec(t,w,ch[i]) : ch[i]; }
  return me;
}
/* SAH-optimal collapse (after Ylitie et al. 2017): F[v][j] = least cost of covering subtree(v) with at most j wide-node
 * children; a child that is an internal binary node becomes a wide node of cost area + the best cover of its two sides
 * with k children in total, a leaf child costs c_leaf * area */
static double g_cleaf=1.0; void lab_set_cleaf(double c){g_cleaf=c;}
typedef struct { double F[9]; int split[9]; double own; int own_split; } dpent;
static void dp_rec(const btree *t,int v,int k,dpent *D){ /* v >= 0 */
  int l=t->left[v], r=t->right[v]; if(l>=0)dp_rec(t,l,k,D); if(r>=0)dp_rec(t,r,k,D);
  double Fl[9],Fr[9]; for(int j=1;j<=k;j++){ Fl[j]= l>=0? D[l].F[j] : g_cleaf*box_area(&t->pbox[~l]); Fr[j]= r>=0? D[r].F[j] : g_cleaf*box_area(&t->pbox[~r]); }
  double best=INFINITY; int bs=1; for(int i=1;i<k;i++){ double c=Fl[i]+Fr[k-i]; if(c<best){best=c;bs=i;} }
  D[v].own=box_area(&t->box[v])+best; D[v].own_split=bs;
  D[v].F[1]=D[v].own; D[v].split[1]=0;
  for(int j=2;j<=k;j++){ double b=D[v].F[j-1]; int sp=D[v].split[j-1]; /* at most j: not worse than at most j-1 */
    for(int i=1;i<j;i++){ double c=Fl[i]+Fr[j-i]; if(c<b){b=c;sp=i+100*j;} } D[v].F[j]=b; D[v].split[j]=sp; } }
static void dp_cut(const btree *t,const dpent *D,int v,int j,int *out,int *n){ /* children covering subtree(v) with at most j roots */
  if(v<0){out[(*n)++]=v;return;}
  int sp=D[v].split[j]; if(sp==0){out[(*n)++]=v;return;} int jj=sp/100,i=sp%100; dp_cut(t,D,t->left[v],i,out,n); dp_cut(t,D,t->right[v],jj-i,out,n); }
static int collapse_dp_rec(const btree *t,const dpent *D,wtree *w,int bnode){
  int me=w->n_nodes++; int k=w->k; int ch[16]; int nc=0; int i=D[bnode].own_split;
  dp_cut(t,D,t->left[bnode],i,ch,&nc); dp_cut(t,D,t->right[bnode],k-i,ch,&nc);
  for(int q=0;q<k;q++){ if(q<nc){ w->cbox[me*k+q]= ch[q]>=0? t->box[ch[q]] : t->pbox[~ch[q]]; } else w->child[me*k+q]=W_EMPTY; }
  for(int q=0;q<nc;q++){ w->child[me*k+q]= ch[q]>=0? collapse_dp_rec(t,D,w,ch[q]) : ch[q]; }
  return me; }
wtree *lab_collapse_dp(const btree *t,int k){ wtree *w=calloc(1,sizeof(wtree)); w->k=k; w->child=malloc(sizeof(int)*t->n*k); w->cbox=malloc(sizeof(box_t)*t->n*k); w->n_nodes=0; w->axis=NULL;
  if(t->n>1){ dpent *D=calloc(t->n,sizeof(dpent)); dp_rec(t,t->root,k,D); collapse_dp_rec(t,D,w,t->root); free(D);} return w; }
wtree *lab_collapse(const btree *t,int k){ wtree *w=calloc(1,sizeof(wtree)); w->k=k; w->child=malloc(sizeof(int)*t->n*k); w->cbox=malloc(sizeof(box_t)*t->n*k); w->n_nodes=0; if(t->n>1)collapse_rec(t,w,t->root); w->axis=NULL; return w; }
/* sort each node's children ascending along the axis with the largest spread of child-box centres (empties last) */
void lab_axis_sort(wtree *w){ int k=w->k; w->axis=malloc(sizeof(int)*w->n_nodes);
  for(int n=0;n<w->n_nodes;n++){ float lo[3]={INFINITY,INFINITY,INFINITY},hi[3]={-INFINITY,-INFINITY,-INFINITY}; int nc=0;
    for(int i=0;i<k;i++) if(w->child[n*k+i]!=W_EMPTY){nc++; for(int a=0;a<3;a++){float c=w->cbox[n*k+i].lo[a]+w->cbox[n*k+i].hi[a]; if(c<lo[a])lo[a]=c; if(c>hi[a])hi[a]=c;}}
    int ax=0; for(int a=1;a<3;a++) if(hi[a]-lo[a]>hi[ax]-lo[ax]) ax=a; w->axis[n]=ax;
    for(int i=1;i<k;i++){ int c=w->child[n*k+i]; box_t b=w->cbox[n*k+i]; if(c==W_EMPTY)continue; float key=b.lo[ax]+b.hi[ax]; int j=i; while(j>0 && (w->child[n*k+j-1]==W_EMPTY || w->cbox[n*k+j-1].lo[ax]+w->cbox[n*k+j-1].hi[ax]>key)){ w->child[n*k+j]=w->child[n*k+j-1]; w->cbox[n*k+j]=w->cbox[n*k+j-1]; j--; } w->child[n*k+j]=c; w->cbox[n*k+j]=b; } } }
int lab_wide_nodes(const wtree *w){return w->n_nodes;}
double lab_wide_fill(const wtree *w){ uint64_t c=0; for(int i=0;i<w->n_nodes*w->k;i++) if(w->child[i]!=W_EMPTY)c++; return (double)c/(w->n_nodes*w->k); }

/* quantise child boxes conservatively to q bits relative to the node's own box (union of children); q=0: exact */
void lab_quantise(wtree *w,int q){ if(q<=0)return; int k=w->k; float levels=(float)((1<<q)-1);
  for(int n=0;n<w->n_nodes;n++){ box_t nb=box_empty(); for(int i=0;i<k;i++) if(w->child[n*k+i]!=W_EMPTY) box_merge(&nb,&w->cbox[n*k+i]);
    for(int a=0;a<3;a++){ float ext=nb.hi[a]-nb.lo[a]; if(!(ext>0))continue; /* power-of-two scale like CWBVH */ int e; frexpf(ext/levels,&e); float sc=ldexpf(1.0f,e);
      for(int i=0;i<k;i++) if(w->child[n*k+i]!=W_EMPTY){ box_t *b=&w->cbox[n*k+i]; float lo=floorf((b->lo[a]-nb.lo[a])/sc),hi=ceilf((b->hi[a]-nb.lo[a])/sc); b->lo[a]=nb.lo[a]+lo*sc; b->hi[a]=nb.lo[a]+hi*sc; } } } }

int g_order_mode=0; void lab_set_order(int m){g_order_mode=m;}
void lab_trace_wide(const wtree *w,const orc_scene *s,const float *rays7,int m,float tmin,int cull_pop,lab_counters *c,int32_t *ids){
  int k=w->k;
  for(int i=0;i<m;i++){ const float *r7=rays7+7*i; ray_t r; for(int q=0;q<3;q++){r.o[q]=r7[q];r.d[q]=r7[3+q];r.inv[q]=1.0f/r.d[q];} r.tm=r7[6];
    float best=INFINITY; int bid=-1; int stack[256]; float stt[256]; int sp=0; int cur=0; c->rays++;
    while(cur!=INT32_MIN){
      if(cur>=0){ c->visits++; int hc[16]; float ht[16]; int nh=0;
        for(int j=0;j<k;j++){ int ch=w->child[cur*k+j]; if(ch==W_EMPTY)continue; c->box_tests++; float te; if(slab(&w->cbox[cur*k+j],&r,tmin,best,&te)){ int p=nh++; while(p>0&&ht[p-1]>te){hc[p]=hc[p-1];ht[p]=ht[p-1];p--;} hc[p]=ch;ht[p]=te; } }
        if(nh>2&&g_order_mode==1){ /* nearest first, the rest in slot order */ int n0=hc[0]; float t0=ht[0]; int q=0; int hc2[16]; float ht2[16]; hc2[q]=n0;ht2[q]=t0;q++; for(int j=0;j<k;j++){int ch=w->child[cur*k+j]; if(ch==W_EMPTY||ch==n0)continue; for(int z=1;z<nh;z++) if(hc[z]==ch){hc2[q]=ch;ht2[q]=ht[z];q++;}} for(int z=0;z<nh;z++){hc[z]=hc2[z];ht[z]=ht2[z];} }
        if(g_order_mode==2&&w->axis){ int rev=r.d[w->axis[cur]]<0; int q=0; int hc2[16]; float ht2[16];
          for(int jj=0;jj<k;jj++){ int j=rev?k-1-jj:jj; int ch=w->child[cur*k+j]; if(ch==W_EMPTY)continue; for(int z=0;z<nh;z++) if(hc[z]==ch){hc2[q]=ch;ht2[q]=ht[z];q++;} }
          for(int z=0;z<nh;z++){hc[z]=hc2[z];ht[z]=ht2[z];} }
        if(nh==0){cur=INT32_MIN;} else { for(int j=nh-1;j>=1;j--){stack[sp]=hc[j];stt[sp]=ht[j];sp++;c->pushes++;} if((uint64_t)sp>c->maxstack)c->maxstack=sp; cur=hc[0]; }
      } else { leaf_hit(s,~cur,r7,tmin,&best,&bid,c); cur=INT32_MIN; }
      if(cur==INT32_MIN){ while(sp){ --sp; if(cull_pop==1&&stt[sp]>best)continue; if(cull_pop==2&&stack[sp]<0&&stt[sp]>best)continue; cur=stack[sp]; break; } }
    }
    if(bid>=0)c->hits++; if(ids)ids[i]=bid; }
}
