// Minimal XML DOM reader for scene files (elements, attributes, comments, declarations, self-closing tags,
// the five predefined entities). The reference uses pugixml (ext/pugixml); scene files only need this subset.
#include <cctype>
#include <cstring>

#include "dtof_host.hpp"

namespace dtof_host {

const std::string *XmlNode::attr(const std::string &name) const {
    for (auto &kv : attrs)
        if (kv.first == name)
            return &kv.second;
    return nullptr;
}

namespace {

struct Parser {
    const std::string &s;
    size_t i = 0;
    explicit Parser(const std::string &text) : s(text) {}

    [[noreturn]] void fail(const std::string &what) const {
        size_t line = 1;
        for (size_t k = 0; k < i && k < s.size(); ++k)
            line += s[k] == '\n';
        throw Error("XML parse error (line " + std::to_string(line) + "): " + what);
    }
    void skip_ws() {
        while (i < s.size() && std::isspace((unsigned char) s[i]))
            ++i;
    }
    bool starts(const char *lit) const { return s.compare(i, strlen(lit), lit) == 0; }
    // skips whitespace, comments, processing instructions, doctype; returns false at end of input
    bool skip_misc() {
        for (;;) {
            skip_ws();
            if (i >= s.size())
                return false;
            if (starts("<!--")) {
                size_t e = s.find("-->", i + 4);
                if (e == std::string::npos)
                    fail("unterminated comment");
                i = e + 3;
            } else if (starts("<?")) {
                size_t e = s.find("?>", i + 2);
                if (e == std::string::npos)
                    fail("unterminated processing instruction");
                i = e + 2;
            } else if (starts("<!")) {
                size_t e = s.find('>', i);
                if (e == std::string::npos)
                    fail("unterminated declaration");
                i = e + 1;
            } else {
                return true;
            }
        }
    }
    std::string name() {
        size_t b = i;
        while (i < s.size() && (std::isalnum((unsigned char) s[i]) || s[i] == '_' || s[i] == '-' || s[i] == ':' || s[i] == '.'))
            ++i;
        if (b == i)
            fail("expected a name");
        return s.substr(b, i - b);
    }
    static std::string unescape(const std::string &v) {
        std::string o;
        for (size_t k = 0; k < v.size(); ++k) {
            if (v[k] != '&') {
                o += v[k];
                continue;
            }
            static const std::pair<const char *, char> ents[] = { { "&amp;", '&' }, { "&lt;", '<' }, { "&gt;", '>' },
                                                                  { "&quot;", '"' }, { "&apos;", '\'' } };
            bool done = false;
            for (auto &e : ents)
                if (v.compare(k, strlen(e.first), e.first) == 0) {
                    o += e.second;
                    k += strlen(e.first) - 1;
                    done = true;
                    break;
                }
            if (!done)
                o += v[k];
        }
        return o;
    }
    std::unique_ptr<XmlNode> element() {
        if (s[i] != '<')
            fail("expected '<'");
        ++i;
        auto n = std::make_unique<XmlNode>();
        n->tag = name();
        for (;;) {
            skip_ws();
            if (i >= s.size())
                fail("unterminated tag <" + n->tag + ">");
            if (starts("/>")) {
                i += 2;
                return n;
            }
            if (s[i] == '>') {
                ++i;
                break;
            }
            std::string k = name();
            skip_ws();
            if (i >= s.size() || s[i] != '=')
                fail("expected '=' after attribute " + k);
            ++i;
            skip_ws();
            if (i >= s.size() || (s[i] != '"' && s[i] != '\''))
                fail("expected a quoted attribute value");
            char q = s[i++];
            size_t e = s.find(q, i);
            if (e == std::string::npos)
                fail("unterminated attribute value");
            n->attrs.emplace_back(k, unescape(s.substr(i, e - i)));
            i = e + 1;
        }
        // content
        for (;;) {
            // text content is ignored (scene files carry everything in attributes)
            size_t lt = s.find('<', i);
            if (lt == std::string::npos)
                fail("missing </" + n->tag + ">");
            i = lt;
            if (starts("</")) {
                i += 2;
                std::string close = name();
                if (close != n->tag)
                    fail("mismatched </" + close + ">, expected </" + n->tag + ">");
                skip_ws();
                if (i >= s.size() || s[i] != '>')
                    fail("expected '>'");
                ++i;
                return n;
            }
            if (starts("<!--") || starts("<?") || starts("<!")) {
                skip_misc();
                continue;
            }
            n->children.push_back(element());
        }
    }
};

} // namespace

std::unique_ptr<XmlNode> parse_xml(const std::string &text) {
    Parser p(text);
    if (!p.skip_misc())
        throw Error("XML parse error: empty document");
    auto root = p.element();
    if (p.skip_misc())
        p.fail("content after the root element");
    return root;
}

} // namespace dtof_host
