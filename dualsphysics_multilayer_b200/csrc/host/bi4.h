// bi4.h -- reader/writer for the DualSPHysics "JBinaryData" container (.bi4/.ibi4/.obi4).
//
// On-disk layout follows the reference (format spec only, independent implementation):
//   file header 64 B            S/JBinaryData.h:166-172, S/JBinaryData.cpp:1156-1166 (MakeFileHead)
//   item / values / array records   S/JBinaryData.cpp:748-770 (InValue), :804-848 (InArray),
//                                   :853-899 (InItem), :1000-1005 (SaveValues)
//   type codes                  S/JBinaryData.h:74-79
// PART / case file contents (value names, array names) follow S/JPartDataBi4.cpp:197-371.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include <memory>

namespace dsph {

enum Bi4Type : int {
  BI4_NULL=0, BI4_TEXT=1, BI4_BOOL=2, BI4_CHAR=3, BI4_UCHAR=4, BI4_SHORT=5, BI4_USHORT=6,
  BI4_INT=7, BI4_UINT=8, BI4_LLONG=9, BI4_ULLONG=10, BI4_FLOAT=11, BI4_DOUBLE=12,
  BI4_INT3=20, BI4_UINT3=21, BI4_FLOAT3=22, BI4_DOUBLE3=23
};

size_t Bi4TypeSize(int type);

struct Bi4Value {
  std::string name;
  int type=BI4_NULL;
  std::string text;          // BI4_TEXT
  unsigned char raw[24]={0}; // every other type, packed little endian as on disk
};

struct Bi4Array {
  std::string name;
  bool hide=false;
  int type=BI4_NULL;
  unsigned count=0;
  std::vector<unsigned char> data; // count*Bi4TypeSize(type) bytes
  const void *ext=nullptr; size_t extsize=0;   // AddArrayRef: the caller's buffer is written as is (no copy), it must outlive the save
  const void* DataPtr()const{ return ext? ext: (const void*)data.data(); }
  size_t DataSize()const{ return ext? extsize: data.size(); }
};

struct Bi4Item {
  std::string name;
  bool hide=false, hidevalues=false;
  std::string fmtfloat="%.7E", fmtdouble="%.15E"; // defaults of JBinaryData (S/JBinaryData.cpp:533-534)
  std::vector<Bi4Value> values;
  std::vector<Bi4Array> arrays;
  std::vector<std::unique_ptr<Bi4Item>> items;

  // ---- value setters (replace if the name exists, like JBinaryData::SetvXXX) ----
  Bi4Value& Slot(const std::string &name,int type);
  void SetText(const std::string &n,const std::string &v);
  void SetBool(const std::string &n,bool v);
  void SetInt(const std::string &n,int v);
  void SetUint(const std::string &n,unsigned v);
  void SetUllong(const std::string &n,unsigned long long v);
  void SetLlong(const std::string &n,long long v);
  void SetFloat(const std::string &n,float v);
  void SetDouble(const std::string &n,double v);
  void SetDouble3(const std::string &n,const double v[3]);
  void SetFloat3(const std::string &n,const float v[3]);
  void SetUint3(const std::string &n,const unsigned v[3]);
  // ---- value getters: throw std::runtime_error when missing unless optional ----
  const Bi4Value* Find(const std::string &n)const;
  bool Exists(const std::string &n)const{ return Find(n)!=nullptr; }
  std::string GetText(const std::string &n,bool optional=false,const std::string &def="")const;
  bool GetBool(const std::string &n,bool optional=false,bool def=false)const;
  int GetInt(const std::string &n,bool optional=false,int def=0)const;
  unsigned GetUint(const std::string &n,bool optional=false,unsigned def=0)const;
  unsigned long long GetUllong(const std::string &n,bool optional=false,unsigned long long def=0)const;
  double GetDouble(const std::string &n,bool optional=false,double def=0)const;
  void GetDouble3(const std::string &n,double out[3],bool optional=false)const;
  // ---- arrays / children ----
  Bi4Array& AddArray(const std::string &n,int type,unsigned count,const void *data);
  Bi4Array& AddArrayRef(const std::string &n,int type,unsigned count,const void *data);
  const Bi4Array* FindArray(const std::string &n)const;
  Bi4Item& AddItem(const std::string &n);
  const Bi4Item* FindItem(const std::string &n)const;
};

// Writes `root` as a bi4 file whose file code is root.name (e.g. "JPartDataBi4").
// all=false: hidden arrays / items are left out (JBinaryData::SaveFile(file,memory,all)).
void Bi4Save(const std::string &file,const Bi4Item &root,bool all=true);
// Appends `item` to the list file (created with header + `parent` when missing): JBinaryData::SaveFileListApp.
void Bi4AppendList(const std::string &file,const std::string &filecode,const Bi4Item &parent,const Bi4Item &item,bool all);
// Loads a bi4 file; filecode (if not empty) is validated against the header like CheckHead.
std::unique_ptr<Bi4Item> Bi4Load(const std::string &file,const std::string &filecode);

// ---------------------------------------------------------------------------------------------
// Case / PART helpers (value and array names of JPartDataBi4).
// ---------------------------------------------------------------------------------------------
struct PartHead {
  std::string runcode="00000000", appname="dsph_b200", casename="Case";
  bool data2d=false; double data2dposy=0;
  unsigned piece=0, npiece=1;
  unsigned long long casenp=0, casenfixed=0, casenmoving=0, casenfloat=0, casenfluid=0;
  double caseposmin[3]={0,0,0}, caseposmax[3]={0,0,0};
  bool npdynamic=false, reuseids=false;
  double dp=0,h=0,b=0,rhop0=0,gamma=0,massbound=0,massfluid=0;
  double mapposmin[3]={0,0,0}, mapposmax[3]={0,0,0};
  int perimode=0; double perixinc[3]={0,0,0}, periyinc[3]={0,0,0}, perizinc[3]={0,0,0};
  int axisdiv=0; bool symmetry=false, splitting=false;
};

struct PartData {
  unsigned cpart=0; double timestep=0; unsigned npok=0, nout=0, step=0; double runtime=0;
  double domainmin[3]={0,0,0}, domainmax[3]={0,0,0};
  unsigned long long nptotal=0, idmax=0;
  double symplecticdtpre=0; // restart value (JSph.cpp:2596-2597)
  std::vector<unsigned> idp;
  std::vector<double> pos;   // xyz packed, npok*3 ("Posd"); filled from "Pos" when the file holds floats
  std::vector<float> vel;    // xyz packed
  std::vector<float> rhop;
  bool possingle=false;      // file stored "Pos" (float3) instead of "Posd"
};

void SavePartFile(const std::string &file,const PartHead &head,const PartData &part,bool posdouble=true);
void LoadPartFile(const std::string &file,PartHead &head,PartData &part,unsigned cpart);
std::string PartFileName(unsigned cpart,unsigned piece=0,unsigned npiece=1);

} // namespace dsph
