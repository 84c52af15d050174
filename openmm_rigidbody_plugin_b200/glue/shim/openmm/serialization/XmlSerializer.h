#ifndef OPENMM_XMLSERIALIZER_H_
#define OPENMM_XMLSERIALIZER_H_
// shim, see Vec3.h: writes / reads the node tree as XML the way OpenMM::XmlSerializer lays it out
// (<rootName type="TypeName" prop="..."> <child .../> </rootName>); minimal parser for its own output.
#include "SerializationProxy.h"
#include <istream>
#include <iterator>
#include <ostream>
#include <typeinfo>
namespace OpenMM {
class XmlSerializer {
public:
    template <class T> static void serialize(const T* object, const std::string& rootName, std::ostream& stream) {
        const SerializationProxy& proxy = SerializationProxy::getProxy(typeid(*object));
        SerializationNode node;
        node.setName(rootName);
        proxy.serialize(object, node);
        node.setStringProperty("type", proxy.getTypeName());
        stream << "<?xml version=\"1.0\" ?>\n";
        write(node, stream, 0);
    }
    template <class T> static T* deserialize(std::istream& stream) {
        std::string text((std::istreambuf_iterator<char>(stream)), std::istreambuf_iterator<char>());
        size_t pos = 0;
        if (text.compare(0, 5, "<?xml") == 0) pos = text.find("?>") + 2;
        SerializationNode node;
        parse(text, pos, node);
        const SerializationProxy& proxy = SerializationProxy::getProxy(node.getStringProperty("type"));
        return reinterpret_cast<T*>(proxy.deserialize(node));
    }
private:
    static void write(const SerializationNode& node, std::ostream& s, int depth) {
        s << std::string(depth, '\t') << '<' << node.getName();
        for (std::map<std::string, std::string>::const_iterator it = node.getProperties().begin(); it != node.getProperties().end(); ++it)
            s << ' ' << it->first << "=\"" << it->second << '"';
        if (node.getChildren().empty()) { s << "/>\n"; return; }
        s << ">\n";
        for (size_t i = 0; i < node.getChildren().size(); i++) write(node.getChildren()[i], s, depth + 1);
        s << std::string(depth, '\t') << "</" << node.getName() << ">\n";
    }
    static void skip(const std::string& t, size_t& p) { while (p < t.size() && (t[p] == ' ' || t[p] == '\n' || t[p] == '\t' || t[p] == '\r')) p++; }
    static void parse(const std::string& t, size_t& p, SerializationNode& node) {
        skip(t, p);
        if (t[p] != '<') throw OpenMMException("XmlSerializer: malformed XML");
        size_t e = t.find_first_of(" />", ++p);
        node.setName(t.substr(p, e - p));
        p = e;
        for (;;) {
            skip(t, p);
            if (t[p] == '/') { p = t.find('>', p) + 1; return; }
            if (t[p] == '>') { p++; break; }
            size_t eq = t.find('=', p), q1 = t.find('"', eq), q2 = t.find('"', q1 + 1);
            node.setStringProperty(t.substr(p, eq - p), t.substr(q1 + 1, q2 - q1 - 1));
            p = q2 + 1;
        }
        for (;;) {
            skip(t, p);
            if (t.compare(p, 2, "</") == 0) { p = t.find('>', p) + 1; return; }
            parse(t, p, node.createChildNode(""));
        }
    }
};
}
#endif
