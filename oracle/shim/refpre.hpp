// Pre-included (g++ -include) ahead of every reference translation unit compiled for oracle/_ref.
// It (1) pulls in the few Qt stand-ins the reference reaches through Qt's transitive includes and
// (2) replaces SimulationItem.hpp (reference: SKIRTcore/SimulationItem.hpp:22-139) -- whose find<T>()
// relies on moc-generated meta-objects -- by an equivalent that searches the same ancestors/children
// with dynamic_cast.  The guard macro below makes the reference's own header a no-op.
#ifndef SHIM_REFPRE_HPP
#define SHIM_REFPRE_HPP
#include <climits>
#include <cfloat>
#include <cmath>
#include <typeinfo>
#include <QtGlobal>
#include <QObject>
#include <QString>
#include <QStringList>
#include <QList>
#include <QVarLengthArray>
#include <QHash>
#include <QPair>

#define SIMULATIONITEM_HPP
class SimulationItem : public QObject
{
protected:
    SimulationItem() : _state(Created) {}
public:
    void setup()
    {
        if (_state > Created) return;
        _state = SetupInProgress;
        setupSelfBefore();
        for (int i = 0; i < children().size(); i++)
        {
            SimulationItem* item = dynamic_cast<SimulationItem*>(children()[i]);
            if (item) item->setup();
        }
        setupSelfAfter();
        _state = SetupDone;
    }
protected:
    virtual void setupSelfBefore() {}
    virtual void setupSelfAfter() {}
public:
    template<class T> T* find(bool setup = true) const
    {
        QObject* ancestor = const_cast<SimulationItem*>(this);
        while (ancestor)
        {
            T* item = dynamic_cast<T*>(ancestor);
            if (!item)
                for (int i = 0; i < ancestor->children().size() && !item; i++)
                    item = dynamic_cast<T*>(ancestor->children()[i]);
            if (item) { if (setup) item->setup(); return item; }
            ancestor = ancestor->parent();
        }
        shimNotFound(typeid(T).name());
    }
    template<class T> T* interface()
    {
        QList<SimulationItem*> cands = interfaceCandidates(typeid(T));
        for (int i = 0; i < cands.size(); i++) { T* p = dynamic_cast<T*>(cands[i]); if (p) return p; }
        return 0;
    }
protected:
    virtual QList<SimulationItem*> interfaceCandidates(const std::type_info&) { return QList<SimulationItem*>() << this; }
    enum State { Created = 0, SetupInProgress, SetupDone };
    State _state;
private:
    // throws FatalError like the reference's find<T>() (SimulationItem.hpp:120-127): callers such as
    // FullInstrument::setupSelfBefore (FullInstrument.cpp:29-49) catch exactly that type
    [[noreturn]] static void shimNotFound(const char* name);
};
#include "FatalError.hpp"
inline void SimulationItem::shimNotFound(const char* name)
{ throw FATALERROR(QString("No simulation item of type ") + name + " found in hierarchy"); }
#endif
