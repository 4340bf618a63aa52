// xmlmini.h -- tiny non-validating XML reader (elements + attributes; comments, <?...?> and text are
// skipped).  Enough for DualSPHysics Case.xml (<case><execution>...), which the reference reads with the
// vendored tinyxml through JXml (S/JXml.cpp).  Header-only, no dependencies.
#pragma once
#include <string>
#include <vector>
#include <memory>
#include <stdexcept>
#include <cstdlib>
#include <cstdio>

namespace dsph {

struct XmlNode {
  std::string name;
  std::vector<std::pair<std::string,std::string>> attrs;
  std::vector<std::unique_ptr<XmlNode>> children;

  const XmlNode* Child(const std::string &n)const{
    for(auto &c:children)if(c->name==n)return c.get();
    return nullptr;
  }
  std::vector<const XmlNode*> Children(const std::string &n)const{
    std::vector<const XmlNode*> r;
    for(auto &c:children)if(c->name==n)r.push_back(c.get());
    return r;
  }
  // Path "a.b.c" relative to this node (like JXml::GetNodeSimple).
  const XmlNode* Path(const std::string &path)const{
    const XmlNode *cur=this; size_t p=0;
    while(cur && p<path.size()){
      size_t q=path.find('.',p); if(q==std::string::npos)q=path.size();
      cur=cur->Child(path.substr(p,q-p)); p=q+1;
    }
    return cur;
  }
  bool Has(const std::string &a)const{ for(auto &kv:attrs)if(kv.first==a)return true; return false; }
  std::string Str(const std::string &a,bool optional=false,const std::string &def="")const{
    for(auto &kv:attrs)if(kv.first==a)return kv.second;
    if(!optional)throw std::runtime_error("xml: attribute '"+a+"' is missing in <"+name+">");
    return def;
  }
  double Dbl(const std::string &a,bool optional=false,double def=0)const{
    if(optional && !Has(a))return def;
    const std::string s=Str(a); char *e=nullptr; const double v=strtod(s.c_str(),&e);
    if(e==s.c_str())throw std::runtime_error("xml: attribute '"+a+"' of <"+name+"> is not a number: "+s);
    return v;
  }
  long long Int(const std::string &a,bool optional=false,long long def=0)const{
    if(optional && !Has(a))return def;
    return (long long)Dbl(a);
  }
  bool Bool(const std::string &a,bool optional=false,bool def=false)const{
    if(optional && !Has(a))return def;
    const std::string s=Str(a);
    return (s=="1"||s=="true"||s=="True"||s=="TRUE");
  }
};

class XmlDoc {
 public:
  std::unique_ptr<XmlNode> root; // synthetic node holding the top-level elements
  static XmlDoc LoadFile(const std::string &file){
    FILE *f=fopen(file.c_str(),"rb");
    if(!f)throw std::runtime_error("xml: cannot open file: "+file);
    std::string s; char buf[65536]; size_t n;
    while((n=fread(buf,1,sizeof(buf),f))>0)s.append(buf,n);
    fclose(f);
    return Parse(s);
  }
  static XmlDoc Parse(const std::string &s){
    XmlDoc d; d.root.reset(new XmlNode()); d.root->name="#root";
    size_t p=0; ParseContent(s,p,*d.root,"");
    return d;
  }
 private:
  static void SkipWs(const std::string &s,size_t &p){ while(p<s.size() && (s[p]==' '||s[p]=='\t'||s[p]=='\n'||s[p]=='\r'))p++; }
  static std::string Unescape(const std::string &v){
    std::string r; r.reserve(v.size());
    for(size_t i=0;i<v.size();i++){
      if(v[i]=='&'){
        if(!v.compare(i,4,"&lt;")){ r+='<'; i+=3; }
        else if(!v.compare(i,4,"&gt;")){ r+='>'; i+=3; }
        else if(!v.compare(i,5,"&amp;")){ r+='&'; i+=4; }
        else if(!v.compare(i,6,"&quot;")){ r+='"'; i+=5; }
        else if(!v.compare(i,6,"&apos;")){ r+='\''; i+=5; }
        else r+=v[i];
      }
      else r+=v[i];
    }
    return r;
  }
  static void ParseContent(const std::string &s,size_t &p,XmlNode &parent,const std::string &closing){
    for(;;){
      size_t lt=s.find('<',p);
      if(lt==std::string::npos){
        if(!closing.empty())throw std::runtime_error("xml: missing </"+closing+">");
        p=s.size(); return;
      }
      p=lt;
      if(!s.compare(p,4,"<!--")){ size_t e=s.find("-->",p+4); if(e==std::string::npos)throw std::runtime_error("xml: unterminated comment"); p=e+3; continue; }
      if(!s.compare(p,2,"<?")){ size_t e=s.find("?>",p+2); if(e==std::string::npos)throw std::runtime_error("xml: unterminated declaration"); p=e+2; continue; }
      if(!s.compare(p,2,"<!")){ size_t e=s.find('>',p+2); if(e==std::string::npos)throw std::runtime_error("xml: unterminated <!"); p=e+1; continue; }
      if(!s.compare(p,2,"</")){
        size_t e=s.find('>',p+2); if(e==std::string::npos)throw std::runtime_error("xml: unterminated closing tag");
        std::string nm=s.substr(p+2,e-p-2);
        while(!nm.empty() && (nm.back()==' '||nm.back()=='\t'||nm.back()=='\n'||nm.back()=='\r'))nm.pop_back();
        if(nm!=closing)throw std::runtime_error("xml: unexpected </"+nm+"> (expected </"+closing+">)");
        p=e+1; return;
      }
      // opening tag
      p++;
      size_t q=p; while(q<s.size() && s[q]!=' '&&s[q]!='\t'&&s[q]!='\n'&&s[q]!='\r'&&s[q]!='>'&&s[q]!='/')q++;
      std::unique_ptr<XmlNode> node(new XmlNode()); node->name=s.substr(p,q-p); p=q;
      bool selfclose=false;
      for(;;){
        SkipWs(s,p);
        if(p>=s.size())throw std::runtime_error("xml: unterminated tag <"+node->name+">");
        if(s[p]=='>'){ p++; break; }
        if(s[p]=='/'){ if(p+1<s.size() && s[p+1]=='>'){ p+=2; selfclose=true; break; } throw std::runtime_error("xml: bad '/' in <"+node->name+">"); }
        size_t e=p; while(e<s.size() && s[e]!='='&&s[e]!=' '&&s[e]!='\t'&&s[e]!='\n'&&s[e]!='\r'&&s[e]!='>')e++;
        std::string an=s.substr(p,e-p); p=e; SkipWs(s,p);
        if(p>=s.size()||s[p]!='=')throw std::runtime_error("xml: attribute '"+an+"' without value in <"+node->name+">");
        p++; SkipWs(s,p);
        if(p>=s.size()||(s[p]!='"'&&s[p]!='\''))throw std::runtime_error("xml: attribute '"+an+"' value is not quoted");
        const char qc=s[p++]; size_t ve=s.find(qc,p);
        if(ve==std::string::npos)throw std::runtime_error("xml: unterminated attribute value");
        node->attrs.emplace_back(an,Unescape(s.substr(p,ve-p))); p=ve+1;
      }
      XmlNode *raw=node.get();
      parent.children.push_back(std::move(node));
      if(!selfclose)ParseContent(s,p,*raw,raw->name);
    }
  }
};

} // namespace dsph
