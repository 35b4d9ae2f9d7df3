// oracle/ref_driver/ParameterObject.h  -- TEST INFRASTRUCTURE, not product code.
//
// Minimal stand-in for the un-vendored third-party header `ParameterObject.h` of
// InteractiveComputerGraphics/GenericParameters (pinned a4e2744e by the reference's
// CMakeLists.txt:69-76; fetched by git at configure time, absent from /root/reference and
// unreachable without network).  The reference's hot path only uses it to *register*
// tunables (TimeStepController.cpp:38-73, SimulationModel.cpp:128-268, Simulation.cpp:59-66)
// and to read gravity (TimeStep.cpp:36).  It carries no arithmetic.  This file restates the
// small public surface those call sites need so that the UNMODIFIED reference sources compile
// into oracle/_ref/.  Written from the call sites, not from the GenericParameters sources.
#pragma once
#include <functional>
#include <memory>
#include <string>
#include <vector>

namespace GenParam {

class ParameterBase {
public:
    enum DataTypes { BOOL, DOUBLE, FLOAT, INT32, UINT32, ENUM, VEC_FLOAT, VEC_DOUBLE, STRING };
    ParameterBase(const std::string &name, const std::string &label) : m_name(name), m_label(label) {}
    virtual ~ParameterBase() {}
    void setReadOnly(bool v) { m_readOnly = v; }
    bool getReadOnly() const { return m_readOnly; }
    const std::string &getName() const { return m_name; }
    std::string m_name, m_label, m_group, m_description;
    bool m_readOnly = false;
};

template <typename T> class NumericParameter : public ParameterBase {
public:
    NumericParameter(const std::string &n, const std::string &l, std::function<T()> g, std::function<void(T)> s)
        : ParameterBase(n, l), m_get(g), m_set(s) {}
    void setMinValue(T v) { m_min = v; }
    void setMaxValue(T v) { m_max = v; }
    T getValue() const { return m_get(); }
    void setValue(T v) { if (m_set) m_set(v); }
    std::function<T()> m_get;
    std::function<void(T)> m_set;
    T m_min = T(), m_max = T();
};

using BoolParameter = NumericParameter<bool>;

class EnumParameter : public NumericParameter<int> {
public:
    using NumericParameter<int>::NumericParameter;
    // GenericParameters hands out consecutive ids starting at 0
    void addEnumValue(const std::string &name, int &id) { id = (int)m_names.size(); m_names.push_back(name); }
    std::vector<std::string> m_names;
};

template <typename T> class VectorParameter : public ParameterBase {
public:
    VectorParameter(const std::string &n, const std::string &l, unsigned dim, T *ptr)
        : ParameterBase(n, l), m_dim(dim), m_ptr(ptr) {}
    unsigned m_dim;
    T *m_ptr;
};

class ParameterObject {
public:
    virtual ~ParameterObject() {}
    virtual void initParameters() {}

    template <typename T> int createNumericParameter(const std::string &n, const std::string &l, T *ptr) {
        return add(new NumericParameter<T>(n, l, [ptr]() { return *ptr; }, [ptr](T v) { *ptr = v; }));
    }
    template <typename T>
    int createNumericParameter(const std::string &n, const std::string &l, std::function<T()> g, std::function<void(T)> s) {
        return add(new NumericParameter<T>(n, l, g, s));
    }
    int createBoolParameter(const std::string &n, const std::string &l, bool *ptr) {
        return add(new BoolParameter(n, l, [ptr]() { return *ptr; }, [ptr](bool v) { *ptr = v; }));
    }
    int createBoolParameter(const std::string &n, const std::string &l, std::function<bool()> g, std::function<void(bool)> s) {
        return add(new BoolParameter(n, l, g, s));
    }
    int createEnumParameter(const std::string &n, const std::string &l, int *ptr) {
        return add(new EnumParameter(n, l, [ptr]() { return *ptr; }, [ptr](int v) { *ptr = v; }));
    }
    int createEnumParameter(const std::string &n, const std::string &l, std::function<int()> g, std::function<void(int)> s) {
        return add(new EnumParameter(n, l, g, s));
    }
    template <typename T> int createVectorParameter(const std::string &n, const std::string &l, unsigned dim, T *ptr) {
        return add(new VectorParameter<T>(n, l, dim, ptr));
    }

    void setGroup(int id, const std::string &g) { m_params[id]->m_group = g; }
    void setDescription(int id, const std::string &d) { m_params[id]->m_description = d; }
    ParameterBase *getParameter(int id) { return m_params[id].get(); }
    unsigned numParameters() const { return (unsigned)m_params.size(); }

    template <typename T> T getValue(int id) const { return static_cast<NumericParameter<T> *>(m_params[id].get())->getValue(); }
    template <typename T> void setValue(int id, T v) { static_cast<NumericParameter<T> *>(m_params[id].get())->setValue(v); }
    template <typename T> T *getVecValue(int id) const { return static_cast<VectorParameter<T> *>(m_params[id].get())->m_ptr; }
    template <typename T> void setVecValue(int id, const T *v) {
        auto *p = static_cast<VectorParameter<T> *>(m_params[id].get());
        for (unsigned i = 0; i < p->m_dim; i++) p->m_ptr[i] = v[i];
    }

protected:
    int add(ParameterBase *p) { m_params.emplace_back(p); return (int)m_params.size() - 1; }
    std::vector<std::unique_ptr<ParameterBase>> m_params;
};

} // namespace GenParam
