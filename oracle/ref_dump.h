// ref_dump.h -- TEST INFRASTRUCTURE.  Tiny binary container used by oracle/ref_harness.cpp (and by
// tests/dumpio.py) to move intermediate solver state out of the reference CPU solver.
//   file   = "DSPHDMP1" { record }
//   record = name[32] (zero padded) | dtype u32 | ncomp u32 | count u64 | count*ncomp*sizeof(dtype) bytes
//   dtype  : 0 u8, 1 u16, 2 u32, 3 f32, 4 f64, 5 i32
#pragma once
#include <cstdio>
#include <cstring>
#include <cstdint>
#include <string>
#include <stdexcept>

struct RefDumpWriter {
  FILE *f;
  explicit RefDumpWriter(const std::string &file){
    f=fopen(file.c_str(),"wb");
    if(!f)throw std::runtime_error("cannot open dump file "+file);
    fwrite("DSPHDMP1",1,8,f);
  }
  ~RefDumpWriter(){ if(f)fclose(f); }
  static unsigned Size(unsigned dtype){ static const unsigned s[6]={1,2,4,4,8,4}; return s[dtype]; }
  void Add(const char *name,unsigned dtype,unsigned ncomp,uint64_t count,const void *data){
    char nm[32]; memset(nm,0,32); strncpy(nm,name,31);
    fwrite(nm,1,32,f); fwrite(&dtype,4,1,f); fwrite(&ncomp,4,1,f); fwrite(&count,8,1,f);
    if(count)fwrite(data,size_t(Size(dtype))*ncomp,count,f);
  }
  void U32(const char *n,unsigned v){ Add(n,2,1,1,&v); }
  void F32(const char *n,float v){ Add(n,3,1,1,&v); }
  void F64(const char *n,double v){ Add(n,4,1,1,&v); }
};
