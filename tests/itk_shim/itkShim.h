// itkShim.h — a MINIMAL stand-in for the ITK classes the cuberille adapter touches
// (SURVEY.md §8b "ITK surface the adapter needs").  ITK itself is not installable in this
// environment (no network, no vendored copy); this shim exists only so that
// include/itkCuberilleImageToMeshFilter.h and tests/cpp/cuberille_test01.cxx can be compiled
// and run here with the same source text a real ITK build would see.  It is test
// infrastructure: nothing in the product includes it.
#ifndef itkShim_h
#define itkShim_h

#include <cstddef>
#include <exception>
#include <iostream>
#include <limits>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#define ITK_EXPORT

namespace itk
{

class Indent
{
public:
  Indent( int n = 0 ) : m_N( n ) {}
  Indent GetNextIndent() const { return Indent( m_N + 2 ); }
  int m_N;
};
inline std::ostream & operator<<( std::ostream & os, const Indent & i ) { for ( int k = 0; k < i.m_N; k++ ) os << ' '; return os; }

class ExceptionObject : public std::exception
{
public:
  ExceptionObject( const std::string & file, unsigned int line, const std::string & desc )
    : m_What( file + ":" + std::to_string( line ) + ": " + desc ), m_Description( desc ) {}
  virtual ~ExceptionObject() throw() {}
  virtual const char * what() const throw() { return m_What.c_str(); }
  const char * GetDescription() const { return m_Description.c_str(); }
private:
  std::string m_What, m_Description;
};
inline std::ostream & operator<<( std::ostream & os, const ExceptionObject & e ) { return os << "itk::ExceptionObject: " << e.what(); }

#define itkExceptionMacro( x )                                                     \
  {                                                                                \
  std::ostringstream itkmsg; itkmsg << "itk::ERROR: " << this->GetNameOfClass() x;  \
  throw ::itk::ExceptionObject( __FILE__, __LINE__, itkmsg.str() );                 \
  }

template <typename T> struct NumericTraits
{
  typedef T PrintType;
  static const T One;
  static T max() { return std::numeric_limits<T>::max(); }
};
template <typename T> const T NumericTraits<T>::One = static_cast<T>( 1 );
template <> struct NumericTraits<unsigned char>
{
  typedef int PrintType;
  static const unsigned char One = 1;
  static unsigned char max() { return 255; }
};
template <> struct NumericTraits<signed char>
{
  typedef int PrintType;
  static const signed char One = 1;
  static signed char max() { return 127; }
};

class LightObject
{
public:
  virtual ~LightObject() {}
  virtual void Register() const { ++m_ReferenceCount; }
  virtual void UnRegister() const { if ( --m_ReferenceCount <= 0 ) delete this; }
  virtual const char * GetNameOfClass() const { return "LightObject"; }
protected:
  LightObject() : m_ReferenceCount( 0 ) {}
  mutable int m_ReferenceCount;
};

template <typename T>
class SmartPointer
{
public:
  SmartPointer() : m_P( 0 ) {}
  SmartPointer( T * p ) : m_P( p ) { if ( m_P ) m_P->Register(); }
  SmartPointer( const SmartPointer & o ) : m_P( o.m_P ) { if ( m_P ) m_P->Register(); }
  template <typename U> SmartPointer( const SmartPointer<U> & o ) : m_P( o.GetPointer() ) { if ( m_P ) m_P->Register(); }
  ~SmartPointer() { if ( m_P ) m_P->UnRegister(); }
  SmartPointer & operator=( const SmartPointer & o ) { return *this = o.m_P; }
  SmartPointer & operator=( T * p ) { if ( p ) p->Register(); if ( m_P ) m_P->UnRegister(); m_P = p; return *this; }
  T * operator->() const { return m_P; }
  T & operator*() const { return *m_P; }
  operator T *() const { return m_P; }
  T * GetPointer() const { return m_P; }
  bool IsNull() const { return m_P == 0; }
  bool IsNotNull() const { return m_P != 0; }
private:
  T * m_P;
};

#define itkWarningMacro( x )                                     \
  { std::ostringstream itkmsg; itkmsg << "WARNING: In " __FILE__ ", line " << __LINE__ << "\n" \
      << this->GetNameOfClass() << " (" << this << "): " x << "\n\n"; std::cerr << itkmsg.str(); }

#define itkNewMacro( x )                                         \
  static Pointer New() { Pointer p = new x; return p; }

#define itkTypeMacro( thisClass, superclass )                    \
  virtual const char * GetNameOfClass() const { return #thisClass; }

#define itkSetMacro( name, type )                                \
  virtual void Set##name( const type _arg )                      \
    { if ( this->m_##name != _arg ) { this->m_##name = _arg; this->Modified(); } }
#define itkGetMacro( name, type )                                \
  virtual type Get##name() { return this->m_##name; }
#define itkSetClampMacro( name, type, min, max )                 \
  virtual void Set##name( type _arg )                            \
    {                                                            \
    const type c = ( _arg < (type)( min ) ? (type)( min ) : ( _arg > (type)( max ) ? (type)( max ) : _arg ) ); \
    if ( this->m_##name != c ) { this->m_##name = c; this->Modified(); } \
    }
#define itkBooleanMacro( name )                                  \
  virtual void name##On() { this->Set##name( true ); }           \
  virtual void name##Off() { this->Set##name( false ); }
#define itkSetObjectMacro( name, type )                          \
  virtual void Set##name( type * _arg )                          \
    { if ( this->m_##name != _arg ) { this->m_##name = _arg; this->Modified(); } }
#define itkGetObjectMacro( name, type )                          \
  virtual type * Get##name() { return this->m_##name.GetPointer(); }

class Object : public LightObject
{
public:
  virtual void Modified() const { ++m_MTime; }
  unsigned long GetMTime() const { return m_MTime; }
  virtual void PrintSelf( std::ostream &, Indent ) const {}
  void Print( std::ostream & os ) const { this->PrintSelf( os, Indent() ); }
protected:
  Object() : m_MTime( 1 ) {}
  mutable unsigned long m_MTime;
};

class DataObject : public Object
{
public:
  void DisconnectPipeline() {}
};

class ProcessObject : public Object
{
public:
  virtual void Update()
    {
    if ( m_Inputs.size() < m_NumberOfRequiredInputs || !m_Inputs[0] ) { itkExceptionMacro( << ": input 0 is not set" ); }
    if ( m_UpdateMTime < this->GetMTime() ) { this->GenerateData(); m_UpdateMTime = this->GetMTime(); }
    }
protected:
  ProcessObject() : m_NumberOfRequiredInputs( 0 ), m_UpdateMTime( 0 ) {}
  void SetNumberOfRequiredInputs( unsigned int n ) { m_NumberOfRequiredInputs = n; }
  virtual void SetNthInput( unsigned int i, DataObject * d )
    { if ( m_Inputs.size() <= i ) m_Inputs.resize( i + 1 ); m_Inputs[i] = d; this->Modified(); }
  DataObject * GetNthInput( unsigned int i ) const { return i < m_Inputs.size() ? m_Inputs[i].GetPointer() : 0; }
  virtual void GenerateData() = 0;
  std::vector< SmartPointer<DataObject> > m_Inputs;
  unsigned int m_NumberOfRequiredInputs;
  unsigned long m_UpdateMTime;
};

// ---- Image ------------------------------------------------------------------------------------------------
template <unsigned int D> struct Size  { unsigned long m_V[D]; unsigned long & operator[]( unsigned i ) { return m_V[i]; } unsigned long operator[]( unsigned i ) const { return m_V[i]; } };
template <unsigned int D> struct Index { long m_V[D]; long & operator[]( unsigned i ) { return m_V[i]; } long operator[]( unsigned i ) const { return m_V[i]; } void Fill( long v ) { for ( unsigned i = 0; i < D; i++ ) m_V[i] = v; } };
template <unsigned int D> struct ImageRegion
{
  Index<D> m_Index; Size<D> m_Size;
  const Size<D> & GetSize() const { return m_Size; }
  const Index<D> & GetIndex() const { return m_Index; }
  void SetSize( const Size<D> & s ) { m_Size = s; }
  void SetIndex( const Index<D> & i ) { m_Index = i; }
};
template <typename T, unsigned int D> struct FixedVector { T m_V[D]; T & operator[]( unsigned i ) { return m_V[i]; } const T & operator[]( unsigned i ) const { return m_V[i]; } };
template <unsigned int D> struct DirectionMatrix { double m_M[D][D]; double * operator[]( unsigned i ) { return m_M[i]; } const double * operator[]( unsigned i ) const { return m_M[i]; } };

template <typename TPixel, unsigned int VDim = 3>
class Image : public DataObject
{
public:
  typedef Image Self;
  typedef SmartPointer<Self> Pointer;
  typedef SmartPointer<const Self> ConstPointer;
  typedef TPixel PixelType;
  typedef Size<VDim> SizeType;
  typedef Index<VDim> IndexType;
  typedef ImageRegion<VDim> RegionType;
  typedef double SpacingValueType;
  typedef FixedVector<double, VDim> SpacingType;
  typedef FixedVector<double, VDim> PointType;
  typedef DirectionMatrix<VDim> DirectionType;
  static const unsigned int ImageDimension = VDim;
  itkNewMacro( Self );
  itkTypeMacro( Image, DataObject );
  void SetRegions( const RegionType & r ) { m_Region = r; }
  void Allocate() { size_t n = 1; for ( unsigned i = 0; i < VDim; i++ ) n *= m_Region.GetSize()[i]; m_Buffer.assign( n, TPixel() ); }
  const RegionType & GetBufferedRegion() const { return m_Region; }
  const RegionType & GetLargestPossibleRegion() const { return m_Region; }
  TPixel * GetBufferPointer() { return m_Buffer.empty() ? 0 : &m_Buffer[0]; }
  const TPixel * GetBufferPointer() const { return m_Buffer.empty() ? 0 : &m_Buffer[0]; }
  const SpacingType & GetSpacing() const { return m_Spacing; }
  const PointType & GetOrigin() const { return m_Origin; }
  const DirectionType & GetDirection() const { return m_Direction; }
  void SetSpacing( const SpacingType & s ) { m_Spacing = s; }
  void SetOrigin( const PointType & o ) { m_Origin = o; }
  void SetDirection( const DirectionType & d ) { m_Direction = d; }
protected:
  Image()
    {
    for ( unsigned i = 0; i < VDim; i++ )
      {
      m_Spacing[i] = 1.0; m_Origin[i] = 0.0; m_Region.m_Index[i] = 0; m_Region.m_Size[i] = 0;
      for ( unsigned j = 0; j < VDim; j++ ) m_Direction[i][j] = ( i == j ) ? 1.0 : 0.0;
      }
    }
  RegionType m_Region;
  SpacingType m_Spacing;
  PointType m_Origin;
  DirectionType m_Direction;
  std::vector<TPixel> m_Buffer;
};

// ---- Mesh -------------------------------------------------------------------------------------------------
template <typename TCoord, unsigned int D> struct Point { TCoord m_V[D]; TCoord & operator[]( unsigned i ) { return m_V[i]; } const TCoord & operator[]( unsigned i ) const { return m_V[i]; } };

template <typename TId, typename TElement>
class VectorContainer : public Object
{
public:
  typedef VectorContainer Self;
  typedef SmartPointer<Self> Pointer;
  itkNewMacro( Self );
  void Reserve( TId n ) { if ( m_V.size() < n ) m_V.resize( n ); }
  void SetElement( TId i, const TElement & e ) { m_V[i] = e; }
  void InsertElement( TId i, const TElement & e ) { if ( m_V.size() <= i ) m_V.resize( i + 1 ); m_V[i] = e; }
  const TElement & GetElement( TId i ) const { return m_V[i]; }
  TId Size() const { return m_V.size(); }
  std::vector<TElement> & CastToSTLContainer() { return m_V; }
protected:
  std::vector<TElement> m_V;
};

template <typename TPixel, unsigned int D> struct DefaultStaticMeshTraits
{
  typedef TPixel PixelType;
  typedef float CoordRepType;
  typedef unsigned long PointIdentifier;
  typedef unsigned long CellIdentifier;
  typedef Point<float, D> PointType;
  struct CellTraits { typedef unsigned long PointIdentifier; typedef TPixel CellPixelType; };
};

template <typename TPixel, typename TCellTraits>
class CellInterface
{
public:
  typedef unsigned long PointIdentifier;
  typedef const PointIdentifier * PointIdConstIterator;
  virtual ~CellInterface() {}
  virtual unsigned int GetNumberOfPoints() const = 0;
  virtual void SetPointIds( const PointIdentifier * ids ) = 0;
  virtual PointIdConstIterator PointIdsBegin() const = 0;
  virtual PointIdConstIterator PointIdsEnd() const = 0;
  class CellAutoPointer
  {
  public:
    CellAutoPointer() : m_P( 0 ), m_Owner( false ) {}
    ~CellAutoPointer() { if ( m_Owner ) delete m_P; }
    void TakeOwnership( CellInterface * p ) { if ( m_Owner ) delete m_P; m_P = p; m_Owner = true; }
    CellInterface * ReleaseOwnership() { m_Owner = false; return m_P; }
    CellInterface * operator->() const { return m_P; }
    CellInterface * GetPointer() const { return m_P; }
  private:
    CellAutoPointer( const CellAutoPointer & ); void operator=( const CellAutoPointer & );
    CellInterface * m_P; bool m_Owner;
  };
};

template <typename TCellInterface, unsigned int N>
class FixedCell : public TCellInterface
{
public:
  typedef typename TCellInterface::PointIdentifier PointIdentifier;
  typedef typename TCellInterface::PointIdConstIterator PointIdConstIterator;
  typedef typename TCellInterface::CellAutoPointer CellAutoPointer;
  typedef CellAutoPointer SelfAutoPointer;
  virtual unsigned int GetNumberOfPoints() const { return N; }
  virtual void SetPointIds( const PointIdentifier * ids ) { for ( unsigned i = 0; i < N; i++ ) m_Ids[i] = ids[i]; }
  virtual PointIdConstIterator PointIdsBegin() const { return m_Ids; }
  virtual PointIdConstIterator PointIdsEnd() const { return m_Ids + N; }
private:
  PointIdentifier m_Ids[N];
};
template <typename TCellInterface> class TriangleCell : public FixedCell<TCellInterface, 3> {};
template <typename TCellInterface> class QuadrilateralCell : public FixedCell<TCellInterface, 4> {};

template <typename TPixel, unsigned int VDim = 3, typename TMeshTraits = DefaultStaticMeshTraits<TPixel, VDim> >
class Mesh : public DataObject
{
public:
  typedef Mesh Self;
  typedef SmartPointer<Self> Pointer;
  typedef TMeshTraits MeshTraits;
  typedef typename MeshTraits::PixelType PixelType;
  typedef typename MeshTraits::CellTraits CellTraits;
  typedef typename MeshTraits::PointType PointType;
  typedef typename MeshTraits::PointIdentifier PointIdentifier;
  typedef typename MeshTraits::CellIdentifier CellIdentifier;
  typedef VectorContainer<PointIdentifier, PointType> PointsContainer;
  typedef typename PointsContainer::Pointer PointsContainerPointer;
  typedef CellInterface<PixelType, CellTraits> CellType;
  typedef typename CellType::CellAutoPointer CellAutoPointer;
  typedef std::vector<CellType *> CellsContainer;
  typedef CellsContainer * CellsContainerPointer;
  itkNewMacro( Self );
  itkTypeMacro( Mesh, DataObject );
  PointsContainer * GetPoints() { return m_Points.GetPointer(); }
  CellsContainer * GetCells() { return &m_Cells; }
  void SetCell( CellIdentifier id, CellAutoPointer & c )
    { if ( m_Cells.size() <= id ) m_Cells.resize( id + 1, 0 ); delete m_Cells[id]; m_Cells[id] = c.ReleaseOwnership(); }
  bool GetCell( CellIdentifier id, CellType *& out ) const { if ( id >= m_Cells.size() || !m_Cells[id] ) return false; out = m_Cells[id]; return true; }
  void SetCellData( CellIdentifier id, PixelType v ) { m_CellData[id] = v; }
  bool GetCellData( CellIdentifier id, PixelType * v ) const
    { typename std::map<CellIdentifier, PixelType>::const_iterator it = m_CellData.find( id ); if ( it == m_CellData.end() ) return false; *v = it->second; return true; }
  unsigned long GetNumberOfPoints() const { return m_Points->Size(); }
  unsigned long GetNumberOfCells() const { return m_Cells.size(); }
protected:
  Mesh() { m_Points = PointsContainer::New(); }
  ~Mesh() { for ( size_t i = 0; i < m_Cells.size(); i++ ) delete m_Cells[i]; }
  PointsContainerPointer m_Points;
  CellsContainer m_Cells;
  std::map<CellIdentifier, PixelType> m_CellData;
};

// ---- ImageToMeshFilter ------------------------------------------------------------------------------------
template <typename TInputImage, typename TOutputMesh>
class ImageToMeshFilter : public ProcessObject
{
public:
  typedef ImageToMeshFilter Self;
  typedef SmartPointer<Self> Pointer;
  itkTypeMacro( ImageToMeshFilter, ProcessObject );
  const TInputImage * GetInput( unsigned int i ) { return static_cast<const TInputImage *>( this->GetNthInput( i ) ); }
  TOutputMesh * GetOutput() { return m_Output.GetPointer(); }
protected:
  ImageToMeshFilter() { m_Output = TOutputMesh::New(); }
  virtual void GenerateOutputInformation() {}
  typename TOutputMesh::Pointer m_Output;
};

template <typename TInputImage, typename TCoordRep = double>
class LinearInterpolateImageFunction : public Object
{
public:
  typedef LinearInterpolateImageFunction Self;
  typedef SmartPointer<Self> Pointer;
  typedef double OutputType;
  itkNewMacro( Self );
  itkTypeMacro( LinearInterpolateImageFunction, Object );
};

// ---- types the reference header only names in typedefs (h:163-175) ---------------------------------------------
template <typename TImage> class ConstShapedNeighborhoodIterator {};
template <typename T, unsigned int D> struct CovariantVector { T m_V[D]; };
template <typename TInputImage, typename TOperatorValueType = float, typename TOutputValueType = float>
class GradientImageFilter : public ProcessObject
{
public:
  typedef GradientImageFilter Self;
  typedef SmartPointer<Self> Pointer;
  typedef CovariantVector<TOutputValueType, 3> OutputPixelType;
  typedef Image<OutputPixelType, 3> OutputImageType;
  itkTypeMacro( GradientImageFilter, ProcessObject );
};
template <typename TInputImage, typename TCoordRep = double>
class VectorLinearInterpolateImageFunction : public Object
{
public:
  typedef VectorLinearInterpolateImageFunction Self;
  typedef SmartPointer<Self> Pointer;
  itkTypeMacro( VectorLinearInterpolateImageFunction, Object );
};

} // namespace itk
#endif
