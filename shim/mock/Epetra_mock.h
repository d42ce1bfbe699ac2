// Stand-ins for the Trilinos (Epetra) and deal.II wrapper classes shim/B200Backend.h touches -- test scaffolding only.
// Each class has exactly the members the shim calls, with Epetra's own names and signatures (Epetra_CrsMatrix.h,
// Epetra_BlockMap.h, Epetra_Import.h, Epetra_MultiVector.h of Trilinos 13.0.1; deal.II 9.3 TrilinosWrappers::SparseMatrix /
// MPI::Vector / BlockSparseMatrix), so that the shim compiles unchanged against the real headers.  Storage is plain
// std::vector filled by tests/cpp/shim_check.cpp.
#pragma once
#include <cstddef>
#include <exception>
#include <map>
#include <string>
#include <vector>

class Epetra_BlockMap {
public:
    std::vector<int> gids;                 // local id -> global id
    std::map<int, int> lid_of;             // global id -> local id
    const std::vector<int>* owner_of_gid = nullptr;   // whole-problem ownership (test scaffolding for RemoteIDList)
    int NumMyElements() const { return (int)gids.size(); }
    int GID(int lid) const { return gids[(size_t)lid]; }
    int LID(int gid) const { auto it = lid_of.find(gid); return it == lid_of.end() ? -1 : it->second; }
    bool MyGID(int gid) const { return lid_of.count(gid) != 0; }
    // owner rank (and its local id, unused here) of every listed global id; collective in the real class
    int RemoteIDList(int NumIDs, const int* GIDList, int* PIDList, int* LIDList) const
    {
        for (int i = 0; i < NumIDs; i++) { PIDList[i] = (*owner_of_gid)[(size_t)GIDList[i]]; if (LIDList) LIDList[i] = -1; }
        return 0;
    }
    void finish() { lid_of.clear(); for (size_t i = 0; i < gids.size(); i++) lid_of[gids[i]] = (int)i; }
};
typedef Epetra_BlockMap Epetra_Map;

class Epetra_Import {
public:
    int n_same = 0;
    std::vector<int> remote_lids, export_lids, export_pids;
    int NumSameIDs() const { return n_same; }
    int NumRemoteIDs() const { return (int)remote_lids.size(); }
    int* RemoteLIDs() const { return const_cast<int*>(remote_lids.data()); }
    int NumExportIDs() const { return (int)export_lids.size(); }
    int* ExportLIDs() const { return const_cast<int*>(export_lids.data()); }
    int* ExportPIDs() const { return const_cast<int*>(export_pids.data()); }
};

class Epetra_CrsMatrix {
public:
    Epetra_Map row_map, col_map;
    const Epetra_Import* importer = nullptr;
    std::vector<int> rowptr, indices;      // local column ids (positions in col_map)
    std::vector<double> values;
    int NumMyRows() const { return row_map.NumMyElements(); }
    int NumMyNonzeros() const { return (int)values.size(); }
    int ExtractMyRowView(int MyRow, int& NumEntries, double*& Values, int*& Indices) const
    {
        NumEntries = rowptr[(size_t)MyRow + 1] - rowptr[(size_t)MyRow];
        Values = const_cast<double*>(values.data()) + rowptr[(size_t)MyRow];
        Indices = const_cast<int*>(indices.data()) + rowptr[(size_t)MyRow];
        return 0;
    }
    const Epetra_Map& RowMap() const { return row_map; }
    const Epetra_Map& ColMap() const { return col_map; }
    const Epetra_Map& DomainMap() const { return row_map; }
    const Epetra_Import* Importer() const { return importer; }
};

class Epetra_FEVector {
public:
    std::vector<double> v;
    int ExtractView(double** A, int* MyLDA) const { *A = const_cast<double*>(v.data()); *MyLDA = (int)v.size(); return 0; }
};

namespace dealii { namespace TrilinosWrappers {
class SparseMatrix {
public:
    Epetra_CrsMatrix m;
    const Epetra_CrsMatrix& trilinos_matrix() const { return m; }
};
class BlockSparseMatrix {
public:
    size_t nb = 0;
    std::vector<SparseMatrix> blocks;      // [nb][nb]
    size_t n_block_rows() const { return nb; }
    size_t n_block_cols() const { return nb; }
    const SparseMatrix& block(size_t i, size_t j) const { return blocks[i * nb + j]; }
    SparseMatrix& block(size_t i, size_t j) { return blocks[i * nb + j]; }
};
namespace MPI {
class Vector {
public:
    Epetra_FEVector ev;
    const Epetra_FEVector& trilinos_vector() const { return ev; }
    Epetra_FEVector& trilinos_vector() { return ev; }
    size_t local_size() const { return ev.v.size(); }
};
}}}

namespace natrium {
typedef dealii::TrilinosWrappers::MPI::Vector distributed_vector;
typedef dealii::TrilinosWrappers::BlockSparseMatrix distributed_sparse_block_matrix;

// DistributionFunctions (solver/DistributionFunctions.h:47-300): at(i), getQ(), updateGhosted()
class DistributionFunctions {
public:
    std::vector<distributed_vector> m_f;
    int ghost_updates = 0;
    size_t getQ() const { return m_f.size(); }
    size_t size() const { return m_f.size(); }
    distributed_vector& at(size_t i) { return m_f.at(i); }
    const distributed_vector& at(size_t i) const { return m_f.at(i); }
    void updateGhosted() { ghost_updates++; }
};

// Stencil (stencils/Stencil.h:53-171): the getters the shim reads
class Stencil {
public:
    size_t D = 0, Q = 0;
    std::vector<double> e, w;              // e[Q][D] scaled
    double scaling = 1.0, cs2 = 1.0 / 3.0;
    size_t getD() const { return D; }
    size_t getQ() const { return Q; }
    double getWeight(size_t i) const { return w[i]; }
    struct Direction { const double* p; double operator()(size_t d) const { return p[d]; } };
    Direction getDirection(size_t i) const { return Direction{e.data() + i * D}; }      // numeric_vector in the reference
    double getScaling() const { return scaling; }
    double getSpeedOfSoundSquare() const { return cs2; }
};

class CollisionException : public std::exception {
    std::string msg;
public:
    explicit CollisionException(const char* m) : msg(m) {}
    const char* what() const noexcept override { return msg.c_str(); }
};
}  // namespace natrium
